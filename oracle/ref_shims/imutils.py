"""Stand-in for imutils (absent from this image) used ONLY to run the reference here.
rotate_bound restated from imutils 0.5.x convenience.py (parity unpinned: the reference pins no
imutils version; upstream computes the centre as (w / 2, h / 2))."""
import cv2
import numpy as np


def rotate_bound(image, angle):
    (h, w) = image.shape[:2]
    (cX, cY) = (w / 2, h / 2)
    M = cv2.getRotationMatrix2D((cX, cY), -angle, 1.0)
    cos = np.abs(M[0, 0])
    sin = np.abs(M[0, 1])
    nW = int((h * sin) + (w * cos))
    nH = int((h * cos) + (w * sin))
    M[0, 2] += (nW / 2) - cX
    M[1, 2] += (nH / 2) - cY
    return cv2.warpAffine(image, M, (nW, nH))
