"""Solver constants (/root/reference/src/option.py:2-10)."""
default_hparas = {
    "GRAD_CLIP": 5.0,           # gradient-norm clip threshold
    "PROGRESS_STEP": 100,       # stdout / TensorBoard refresh period (steps)
    "DEV_STEP_RATIO": 1.2,      # greedy-validation decode steps = ratio * longest reference length
    "DEV_N_EXAMPLE": 4,         # examples shown in TensorBoard
    "TB_FLUSH_FREQ": 180,       # seconds
}
