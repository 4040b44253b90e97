"""pelvistim-fem_b200 — B200-native steady-current-conduction FEM engine.

Drop-in for the solve step that pelvistim-fem delegates to ElmerGrid/ElmerSolver
(``step03_ankle_layers/run_layered_sweep.py:1077,1099``).  Host side is Python
(like the reference); all arithmetic on the hot path runs in hand-written
sm_100a CUDA kernels behind the C-ABI of ``include/ptfem.h`` (``libptfem.so``).
"""
__version__ = "0.1.0"
