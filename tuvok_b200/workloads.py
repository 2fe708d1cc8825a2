"""The BASELINE.json configurations as data (no oracle imports: bench.py's GPU arm uses this)."""
import numpy as np

from . import _lib as L
from . import synth
from .tf import TransferFunction1D, TransferFunction2D

# name -> dict.  Cameras: reference defaults (eye (0,0,1.6), fov 50) + a 36-step orbit
# (rotation about Y in 10 degree steps, then 20 degrees about X), SURVEY 8d.
WORKLOADS = {
    # configs[0]: GLRaycaster 1D-TF, 256^3 u8 single brick, 512x512
    "c1": dict(kind=synth.V_SPH, size=(256, 256, 256), dtype=L.U8, brick=260, overlap=2, mode=L.RM_1DTRANS,
               lighting=False, width=512, height=512, tf=(0.3, 0.3), label="256^3 u8 single brick 1D-TF 512x512"),
    # configs[1]: 512^3 u16 bricked 32^3+ghost, 1D TF + ERT, 1024x1024
    "c2": dict(kind=synth.V_NOISE, size=(512, 512, 512), dtype=L.U16, brick=36, overlap=2, mode=L.RM_1DTRANS,
               lighting=False, width=1024, height=1024, tf=(0.3, 0.4),
               label="512^3 u16 bricked 36^3 1D-TF 1024x1024"),
    # configs[2]: 2048^3 u16 GridLeaper, 2D TF + gradient lighting, 1920x1080 (the headline)
    "c3": dict(kind=synth.V_NOISE, size=(2048, 2048, 2048), dtype=L.U16, brick=36, overlap=2, mode=L.RM_2DTRANS,
               lighting=True, width=1920, height=1080, tf=(0.3, 0.4),
               label="2048^3 u16 bricked 36^3 GridLeaper 2D-TF+lighting 1920x1080"),
    # second reported workload: C3 with a TRANSLUCENT table (alpha_max 2/255 instead of 16/255): rays
    # are not cut short by early termination, so frames/s and Gsamples/s are shown where rays are long
    "c3t": dict(kind=synth.V_NOISE, size=(2048, 2048, 2048), dtype=L.U16, brick=36, overlap=2, mode=L.RM_2DTRANS,
                lighting=True, width=1920, height=1080, tf=(0.3, 0.4), alpha_max=2,
                label="2048^3 u16 bricked 36^3 GridLeaper 2D-TF (translucent, alpha_max 2/255)+lighting 1920x1080"),
    # configs[3]: 1024^3 f32 isosurface + lighting
    "c4": dict(kind=synth.V_SPH, size=(1024, 1024, 1024), dtype=L.F32, brick=36, overlap=2, mode=L.RM_ISOSURFACE,
               lighting=True, width=1920, height=1080, tf=(0.3, 0.4), iso=0.35,
               label="1024^3 f32 bricked 36^3 isosurface+lighting 1920x1080"),
    # configs[4]: 8192^3 u8 (512 GiB), 128^3 bricks, multi-resolution LOD out-of-core.  The volume exists nowhere at once:
    # it is the procedural dataset of tvk_set_procedural_volume (bricks generated on demand by host threads into pinned
    # staging memory and streamed to the pool); sort-last shards the bricks across the ranks by construction (a rank only
    # ever asks for the bricks of its block)
    "c5": dict(kind=synth.V_NOISE, size=(8192, 8192, 8192), dtype=L.U8, brick=128, overlap=2, mode=L.RM_1DTRANS,
               lighting=True, width=1920, height=1080, tf=(0.3, 0.4), procedural=True,
               # the camera flies INSIDE the volume (the volume is moved towards the eye and turned about its centre), so the
               # bricks near the eye are needed at the finest levels and every view needs other ones; a translucent table
               # (alpha x 1/16) keeps rays alive through several LoD shells
               translate=(0.0, 0.0, 1.2), alpha_scale=1.0 / 16.0,
               label="8192^3 u8 (512 GiB) procedural, bricked 128^3, LOD out-of-core fly-through, 1D-TF+lighting 1920x1080"),
}


# Samples (ComputeColorFromVolume / GetVolumeHit evaluations) of one converged frame, mean over the 36-view orbit, as
# counted on the device by the counting variant of the traversal kernel (bench.py reports the live count in
# config.samples_per_frame; these are the values of profiles/r1d_bench_*.json).  The workloads are seeded and the
# arithmetic is fixed, so this is a constant of the workload; the CPU reference arm (bench.py --impl reference), which
# cannot render the full frame in bounded time, converts its measured samples/s to frames/s with it.
SAMPLES_PER_FRAME = {"c2": 44081668.6, "c3": 99164900.1, "c4": 19898366.8}


def transfer_functions(w):
    """(TransferFunction1D, TransferFunction2D) of a workload: SetStdFunction ramp on 4096 (256 for u8)
    entries; 2D TF = one rectangular swatch, as wide as the 1D table, 256 gradient bins."""
    n = 256 if w["dtype"] == L.U8 else 4096
    t1 = TransferFunction1D(n)
    t1.SetStdFunction(*w["tf"])
    if "alpha_scale" in w:
        c = t1.color.copy()
        c[:, 3] *= np.float32(w["alpha_scale"])
        t1.Set(c)
    t2 = TransferFunction2D.rectangle(w=n, h=256, x0=0.02, x1=0.9, alpha_max=w.get("alpha_max", 16))
    return t1, t2


def orbit_rotation(step, n_steps=36):
    from .renderer import rotation_x, rotation_y
    import numpy as np
    return (rotation_y(360.0 * step / n_steps) @ rotation_x(20.0)).astype(np.float32)
