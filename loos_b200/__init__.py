"""loos_b200 -- B200-native all-to-all optimal-superposition RMSD (the hot path
behind LOOS's rmsds / multi-rmsds / rmsd2ref).

Python here is plumbing only: it mirrors the reference's L3 function boundary
(readCoords -> centeredRMSD matrix; superposition/rmsd per frame) on top of the
C ABI in include/loosb200.h.  All arithmetic runs in hand-written sm_100a CUDA
kernels (loos_b200/csrc); there is no CPU path in this package.
"""
from .api import (FrameSet, rmsds, rmsds_cross, rmsd2ref, rmsd2ref_iterative, part_rows,  # noqa: F401
                  rmsds_align, rmsds_multi, rmsds_cross_multi, rmsds_text, rmsds_text_multi, PinnedMatrix, stage_multi,
                  LoosB200Error)

__all__ = ["FrameSet", "rmsds", "rmsds_cross", "rmsd2ref", "rmsd2ref_iterative", "part_rows",
           "rmsds_align", "rmsds_multi", "rmsds_cross_multi", "rmsds_text", "rmsds_text_multi", "PinnedMatrix", "stage_multi",
           "LoosB200Error"]
