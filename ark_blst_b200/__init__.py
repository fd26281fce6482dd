"""ark-blst_b200 — B200-native BLS12-381 MSM engine behind ark-blst's arkworks surface.

Only what the hot path needs: csrc/ (sm_100a kernels + the C-ABI of include/b200msm.h) and this
thin host-side mirror of the reference's `VariableBaseMSM` impls.
"""
from .msm import (  # noqa: F401
    G1,
    G2,
    G1Projective,
    G2Projective,
    MsmError,
    ResidentBases,
    imad_peak,
    last_phase_ms,
    last_plan,
    run_device,
    run_table_device,
    table_build_device,
    table_plan,
    sum_partials_device,
    synth_bases_device,
    synth_scalars_device,
)
