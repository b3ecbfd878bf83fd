// GPU-side PNG image data (csrc/rr_png_gpu.cu): buffers of one encoding pass over F streams.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct rr_png_bufs {
    int W, H;                      // image size; a stream is H scanlines of 1 + 4 W bytes (filter byte + RGBA)
    size_t n, n_pad;               // filtered bytes per stream, and the same rounded up to 128
    size_t cap;                    // bytes reserved per output stream (a multiple of 4)
    int nchunks;                   // ceil(n / 128)
    uint8_t *filt;                 // [F][n_pad]   Sub-filtered scanlines
    unsigned *hist;                // [F][256]
    unsigned long long *adler;     // [F][2]       sum of bytes, sum of (n - i) * byte
    unsigned *codes;               // [F][257]     bit-reversed code | length << 16
    unsigned *chunk_bits;          // [F][nchunks] bits per chunk, then (k_png_scan) the chunk's first bit in the stream
    uint8_t *stream;               // [F][cap]     complete zlib streams
    unsigned *sizes;               // [F]          bytes of each stream (0: did not fit -- cannot happen with cap = 1.25 n)
};

cudaError_t rr_png_upload_constants();
// src: [F][H][W][3] uint8 BGR (mask == false) or [F][H][W] uint8 colormap indices (mask == true), device memory
cudaError_t rr_launch_png_encode(const rr_png_bufs &p, const uint8_t *src, bool mask, int F, cudaStream_t st);
