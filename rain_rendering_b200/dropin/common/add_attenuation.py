"""common.add_attenuation.FogRain (common/add_attenuation.py:26-95) on the GPU: fog_rain_layer
runs the fused fog kernel (csrc/rr_kernels.cu: k_fog) through rr_fog_only."""
import numpy as np


class FogRain:
    def __init__(self, rain_intensity, focal, f_number, angle, exposure=2, camera_gain=20):
        self.rain_intensity = rain_intensity
        self.angle = angle                  # the reference only ever passes 90 (generator.py:232)
        self.focal = focal
        self.f_number = f_number
        self.exposure_time = exposure * 1e-3
        self.exposure_ms = exposure
        self.camera_gain = camera_gain
        self._ctx = None
        self._shape = None

    def calc_beta_ext(self):
        return 0.312 * self.rain_intensity ** 0.67

    def fog_rain_layer(self, image, depth):
        """image: (H,W,3) float64 BGR in [0,1] (cv2.imread(...)/255.0); depth: (H,W) metres.
        Returns (H,W,3) float64.  The image must be exactly representable as uint8/255."""
        from rain_rendering_b200.api import RainContext
        assert self.angle == 90, "only the reference's angle=90 branch is implemented"
        u8 = np.rint(np.asarray(image) * 255.0).astype(np.uint8)
        assert np.array_equal(u8 / 255.0, image), "fog_rain_layer expects an 8-bit image scaled to [0,1]"
        H, W = u8.shape[:2]
        if self._ctx is None or self._shape != (H, W):
            self._ctx = RainContext(0)
            self._ctx.set_camera(W, H, focal_mm=self.focal * 1000., f_number=self.f_number, exposure_ms=self.exposure_ms,
                                 gain=self.camera_gain, fallrate=self.rain_intensity, max_batch=1)
            self._shape = (H, W)
        out = self._ctx.fog_only(np.ascontiguousarray(u8[None]), np.ascontiguousarray(depth[None], dtype=np.float32))
        return np.ascontiguousarray(np.moveaxis(out[0], 0, -1))
