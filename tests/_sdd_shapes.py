"""Initial shapes of pf/sdd_test.go (TestClassicalNucleation), shared by the oracle and GPU tests."""
import numpy as np


def box_blur_5x5(data: np.ndarray, n: int) -> np.ndarray:
    """pfutil.Blur with BoxKernel{Width: 2} on an n x n periodic grid (pfutil/blur.go:14-70): Cutoff()
    = 2, so the window is the 5 x 5 box around each node and every weight is 1."""
    v = data.reshape(n, n)
    acc = np.zeros_like(v)
    for dr in range(-2, 3):
        for dc in range(-2, 3):
            acc += np.roll(np.roll(v, dr, axis=0), dc, axis=1)
    return (acc / 25.0).reshape(-1)


def insert_circle_at_center(data: np.ndarray, n: int, radius: int):
    # pf/sdd_test.go:151-161
    i = np.arange(n * n)
    dx = i // n - n // 2
    dy = i % n - n // 2
    data[dx * dx + dy * dy <= radius * radius] = 1.0
