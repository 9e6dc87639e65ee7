"""lehamoc_b200 -- B200-native (sm_100a, float64) implementation of LeHaMoC's leptonic per-timestep kinetic
update behind the reference's own Python surface:

    lehamoc_b200.Code.LeMoC(params, fileName)         <-> Fit_emcee/Code.py::LeMoC
    python -m lehamoc_b200.LeMoC Parameters.txt out   <-> python LeMoC.py Parameters.txt out
    lehamoc_b200.LeHaMoC_f.*                          <-> LeMoC/LeHaMoC_f.py (function-level twins)

The compute path is liblhm.so (lehamoc_b200/csrc, C ABI in include/lhm.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
