"""The named seeded scenes behind tests/golden/*.npz (shared by the generator, the CPU and the GPU tests)."""
import numpy as np

import tuvok_b200 as tb
from oracle import orc
from scene import Scene
from tuvok_b200 import synth

ROT = (tb.rotation_y(30.0) @ tb.rotation_x(20.0)).astype(np.float32)

SCENES = {
    # BASELINE configs[0] in miniature: single brick, u8, 1D TF, no lighting
    "c1_single_brick_1d": dict(kind=synth.V_SPH, size=(48, 48, 48), dtype=orc.U8, brick=52, overlap=2, width=64, height=64,
                               tf_center=0.3, tf_inv_gradient=0.3),
    # configs[1] in miniature: u16, 36^3 bricks (the baked-size kernel), 1D TF with early ray termination
    "c2_bricked36_1d_ert": dict(kind=synth.V_NOISE, size=(96, 96, 96), dtype=orc.U16, brick=36, overlap=2, width=80,
                                height=80, rotation=ROT, tf_center=0.2, tf_inv_gradient=0.2),
    # configs[2] in miniature: GridLeaper page-table traversal, 2D TF + gradient lighting
    "c3_bricked36_2d_lit": dict(kind=synth.V_NOISE, size=(96, 80, 72), dtype=orc.U16, brick=36, overlap=2, width=96,
                                height=64, rotation=ROT, mode=orc.RM_2DTRANS, lighting=True),
    # configs[3] in miniature: f32 isosurface with lighting and min/max empty-space skipping
    "c4_f32_iso": dict(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.F32, brick=20, overlap=2, width=64, height=64,
                       rotation=ROT, mode=orc.RM_ISOSURFACE, lighting=True, isovalue=0.35),
    # 1D TF + lighting, u8, odd volume / brick sizes (ragged last bricks)
    "ragged_1d_lit": dict(kind=synth.V_NOISE, size=(70, 45, 58), dtype=orc.U8, brick=20, overlap=2, width=72, height=56,
                          rotation=ROT, lighting=True, tf_center=0.25, tf_inv_gradient=0.3),
    # camera inside the volume (near-plane ray entry), anisotropic voxels
    "inside_aniso_2d": dict(kind=synth.V_NOISE, size=(64, 64, 32), dtype=orc.U16, brick=20, overlap=2, width=64, height=48,
                            rotation=ROT, translation=tb.translation(0.05, 0.0, 1.25), mode=orc.RM_2DTRANS,
                            scale=(1.0, 1.0, 2.0)),
}


def make(name, **over):
    kw = dict(SCENES[name])
    kw.update(over)
    return Scene(**kw)
