"""Mask coefficient tables of the reference's sample operators (workload parameters).

These are the numeric parameters the BASELINE.json configs are quoted on; each table cites
the sample that defines it (paths relative to the Hipacc tree, samples-public/).
"""
import numpy as np

# 1_Local_Operators/Gaussian_Blur/src/main.cpp:96-118
GAUSS3 = np.array([[0.057118, 0.124758, 0.057118],
                   [0.124758, 0.272496, 0.124758],
                   [0.057118, 0.124758, 0.057118]], dtype=np.float32)
GAUSS5 = np.array([[0.005008, 0.017300, 0.026151, 0.017300, 0.005008],
                   [0.017300, 0.059761, 0.090339, 0.059761, 0.017300],
                   [0.026151, 0.090339, 0.136565, 0.090339, 0.026151],
                   [0.017300, 0.059761, 0.090339, 0.059761, 0.017300],
                   [0.005008, 0.017300, 0.026151, 0.017300, 0.005008]], dtype=np.float32)
GAUSS7 = np.array([[0.000841, 0.003010, 0.006471, 0.008351, 0.006471, 0.003010, 0.000841],
                   [0.003010, 0.010778, 0.023169, 0.029902, 0.023169, 0.010778, 0.003010],
                   [0.006471, 0.023169, 0.049806, 0.064280, 0.049806, 0.023169, 0.006471],
                   [0.008351, 0.029902, 0.064280, 0.082959, 0.064280, 0.029902, 0.008351],
                   [0.006471, 0.023169, 0.049806, 0.064280, 0.049806, 0.023169, 0.006471],
                   [0.003010, 0.010778, 0.023169, 0.029902, 0.023169, 0.010778, 0.003010],
                   [0.000841, 0.003010, 0.006471, 0.008351, 0.006471, 0.003010, 0.000841]], dtype=np.float32)
GAUSS = {3: GAUSS3, 5: GAUSS5, 7: GAUSS7}

# 3_Preprocessing/Sobel/src/main.cpp:128-178
SOBEL3_X = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=np.int32)
SOBEL3_Y = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], dtype=np.int32)
SOBEL5_X = np.array([[-1, -2, 0, 2, 1], [-4, -8, 0, 8, 4], [-6, -12, 0, 12, 6],
                     [-4, -8, 0, 8, 4], [-1, -2, 0, 2, 1]], dtype=np.int32)
SOBEL5_Y = np.array([[-1, -4, -6, -4, -1], [-2, -8, -12, -8, -2], [0, 0, 0, 0, 0],
                     [2, 8, 12, 8, 2], [1, 4, 6, 4, 1]], dtype=np.int32)
SOBEL_NORM = {3: 4, 5: 48, 7: 640}

# 1_Local_Operators/Laplace/src/main.cpp:106-124 (SIZE_X == 1, 3, 5 variants)
LAPLACE3_4N = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], dtype=np.int32)
LAPLACE3 = np.array([[2, 0, 2], [0, -8, 0], [2, 0, 2]], dtype=np.int32)
LAPLACE5 = np.ones((5, 5), dtype=np.int32)
LAPLACE5[2, 2] = -24

# 3_Preprocessing/Harris_Corner/src/main.cpp:186-226
HARRIS_GAUSS3 = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.int32)
HARRIS_DX = np.array([[-1, 0, 1], [-1, 0, 1], [-1, 0, 1]], dtype=np.int32)
HARRIS_DY = np.array([[-1, -1, -1], [0, 0, 0], [1, 1, 1]], dtype=np.int32)
HARRIS_K = 0.04
HARRIS_THRESHOLD = 20000.0


def bilateral_mask(size):
    """3_Preprocessing/Bilateral_Filter/src/main.cpp:106-140: the sample's closeness tables are
    exp(-(dx^2+dy^2)/(2*s^2)) printed to 6 decimals, with 2*s^2 chosen so the corner tap is e^-4."""
    h = size // 2
    d = np.arange(-h, h + 1, dtype=np.float64)
    r2 = d[None, :] ** 2 + d[:, None] ** 2
    scale = 4.0 / (2.0 * h * h)
    return np.round(np.exp(-r2 * scale), 6).astype(np.float32)
