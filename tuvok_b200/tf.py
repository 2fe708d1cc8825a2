"""Host-side mirrors of Tuvok's transfer-function classes (only what the renderer consumes).

TransferFunction1D: IO/TransferFunction1D.cpp:74-113 (SetStdFunction), :311-330 (GetByteArray,
truncating float->u8), :362-371 (ComputeNonZeroLimits on the float alpha).
TransferFunction2D: the reference rasterises swatch polygons with Qt (TransferFunction2D.cpp:294-361),
which is not available; the table is supplied pre-rasterised as RGBA8 and only
ComputeNonZeroLimits (:378-397) is mirrored.
"""
import numpy as np


class TransferFunction1D:
    def __init__(self, size=256):
        self.color = np.zeros((int(size), 4), np.float32)
        self.value_bbox = (int(size), 0)
        self.SetStdFunction(0.5, 0.5)

    def GetSize(self):
        return self.color.shape[0]

    def Set(self, rgba):
        self.color = np.ascontiguousarray(rgba, np.float32).reshape(-1, 4).copy()
        self.ComputeNonZeroLimits()

    def SetStdFunction(self, center=0.5, inv_gradient=0.5):
        n = self.color.shape[0]
        c = np.float32(min(max(0.0, center), 1.0))
        g = np.float32(min(max(0.0, inv_gradient), 1.0))
        ic = int(np.float32(n - 1) * c)
        ig = int(np.float32(n - 1) * g)
        start = 0 if ig // 2 > ic else ic - ig // 2
        end = n if ig // 2 + ic > n else ic + ig // 2
        i = np.arange(start, end, dtype=np.int64)
        x = (i - ic + ig // 2).astype(np.float32) / np.float32(ig) if ig else np.zeros(len(i), np.float32)
        ramp = np.float32(3) * x * x - np.float32(2) * x * x * x
        for comp in range(4):
            self.color[:start, comp] = 0
            self.color[start:end, comp] = ramp
            self.color[end:, comp] = 1
        self.ComputeNonZeroLimits()

    def GetByteArray(self, used_range=255):
        v = np.maximum(np.float32(0), np.minimum(self.color, np.float32(1))) * np.float32(used_range)
        return v.astype(np.uint8)   # C cast: truncation

    def ComputeNonZeroLimits(self):
        nz = np.nonzero(self.color[:, 3] != 0)[0]
        self.value_bbox = (int(nz[0]), int(nz[-1])) if len(nz) else (self.color.shape[0], 0)

    def GetNonZeroLimits(self):
        return self.value_bbox


class TransferFunction2D:
    def __init__(self, rgba8):
        """rgba8: uint8 array [h, w, 4] (row = gradient bin, column = value bin)."""
        self.pixels = np.ascontiguousarray(rgba8, np.uint8)
        assert self.pixels.ndim == 3 and self.pixels.shape[2] == 4
        self.ComputeNonZeroLimits()

    def GetSize(self):
        return self.pixels.shape[1], self.pixels.shape[0]

    def GetByteArray(self):
        return self.pixels

    def ComputeNonZeroLimits(self):
        h, w = self.pixels.shape[:2]
        ys, xs = np.nonzero(self.pixels[:, :, 3] != 0)
        self.value_bbox = (int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())) if len(xs) else (w, 0, h, 0)

    def GetNonZeroLimits(self):
        return self.value_bbox

    @staticmethod
    def rectangle(w=256, h=256, x0=0.25, x1=0.9, y0=0.0, y1=1.0, color=(255, 160, 64), alpha_max=96):
        """Deterministic stand-in for a rectangular swatch: alpha ramps up linearly in x."""
        px = np.zeros((h, w, 4), np.uint8)
        xa, xb = int(x0 * w), int(x1 * w)
        ya, yb = int(y0 * h), int(y1 * h)
        ramp = (np.arange(xa, xb) - xa + 1) * alpha_max // max(1, xb - xa)
        px[ya:yb, xa:xb, 0] = color[0]
        px[ya:yb, xa:xb, 1] = (color[1] * (np.arange(xa, xb) - xa) // max(1, xb - xa)).astype(np.uint8)[None, :]
        px[ya:yb, xa:xb, 2] = color[2]
        px[ya:yb, xa:xb, 3] = ramp.astype(np.uint8)[None, :]
        return TransferFunction2D(px)
