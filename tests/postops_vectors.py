"""Known-answer vectors the reference holds for Filters.median / Filters.wiener / PeakFinding.argrel*
(re-encoded by hand from the cited lines; shared by the oracle test and the GPU parity test)."""
import numpy as np

# test/nx_signal/filters_test.exs:6-12
MEDIAN_1D = (np.array([10, 9, 8, 7, 1, 4, 5, 3, 2, 6], dtype=np.int32), (3,),
             np.array([9.0, 8.0, 7.0, 4.0, 4.0, 4.0, 3.0, 3.0, 3.0, 3.0], dtype=np.float32))
# test/nx_signal/filters_test.exs:14-32
_T2 = np.array([[31, 11, 17, 13, 1], [1, 3, 19, 23, 29], [19, 5, 7, 37, 2]], dtype=np.int32)
_E2 = np.tile(np.array([11.0, 13.0, 17.0, 17.0, 17.0], dtype=np.float32), (3, 1))
MEDIAN_2D = (_T2, (3, 3), _E2)
# test/nx_signal/filters_test.exs:34-97
_T3 = np.array([
    [[31, 11, 17, 13, 1], [1, 3, 19, 23, 29], [19, 5, 7, 37, 2]],
    [[19, 5, 7, 37, 2], [1, 3, 19, 23, 29], [31, 11, 17, 13, 1]],
    [[1, 3, 19, 23, 29], [31, 11, 17, 13, 1], [19, 5, 7, 37, 2]]], dtype=np.int32)
MEDIAN_3D_K331 = (_T3, (3, 3, 1), np.tile(np.array([19.0, 5.0, 17.0, 23.0, 2.0], dtype=np.float32), (3, 3, 1)))
MEDIAN_3D_K333 = (_T3, (3, 3, 3), np.tile(np.array([11.0, 13.0, 17.0, 17.0, 17.0], dtype=np.float32), (3, 3, 1)))

# test/nx_signal/filters_test.exs:121-177 (noise estimated) and :179-244 (noise given)
WIENER_IM = np.arange(1.0, 16.0, dtype=np.float64).reshape(3, 5)
WIENER_EST_F64 = np.array([
    [1.7777777777777777, 3.0, 3.6666666666666665, 4.333333333333333, 3.111111111111111],
    [4.3366520642506305, 7.0, 8.0, 9.0, 7.58637597408283],
    [4.692197051420351, 7.261706150595039, 8.748939779474131, 10.157992415073023, 9.813815742524799]], dtype=np.float64)
WIENER_EST_F32 = np.array([
    [1.7777777910232544, 3.0, 3.6666667461395264, 4.333333492279053, 3.1111111640930176],
    [4.3366522789001465, 7.0, 8.0, 9.0, 7.586376190185547],
    [4.692196846008301, 7.261706352233887, 8.748939514160156, 10.157992362976074, 9.81381607055664]], dtype=np.float32)
WIENER_N10_F64 = np.array([
    [1.7777777777777777, 3.0, 3.5882352941176467, 4.238095238095238, 3.7397034596375622],
    [5.193548387096774, 7.0, 8.0, 9.0, 8.829787234042554],
    [7.941747572815534, 9.702702702702702, 10.938931297709924, 12.137254901960784, 12.485549132947977]], dtype=np.float64)
WIENER_N10_F32 = np.array([
    [1.7777777910232544, 3.0, 3.588235378265381, 4.238095283508301, 3.739703416824341],
    [5.193548202514648, 7.0, 8.0, 9.0, 8.829787254333496],
    [7.941747665405273, 9.702702522277832, 10.938931465148926, 12.13725471496582, 12.485548973083496]], dtype=np.float32)
# lib/nx_signal/filters.ex:68-78 (doctest): kernel {2, 2}, noise 10
WIENER_DOC = (np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]], dtype=np.float32), (2, 2), 10,
              np.array([[0.25, 0.75, 1.25], [1.25, 3.0, 4.0], [2.75, 6.0, 7.0]], dtype=np.float32))

# lib/nx_signal/peak_finding.ex:36-128 (argrelmin), :158-250 (argrelmax), :296-331 (argrelextrema)
PEAK_X1 = np.array([2, 1, 2, 3, 2, 0, 1, 0], dtype=np.int32)
PEAK_X2 = np.array([[1, 2, 1, 2], [6, 2, 0, 0], [5, 3, 4, 4]], dtype=np.int32)
# (function, data, kwargs, leading valid rows, valid count)
PEAKS = [
    ("less", PEAK_X1, dict(), [[1], [5]], 2),
    ("less", PEAK_X1, dict(order=3), [[1]], 1),
    ("less", PEAK_X2, dict(), [[1, 2], [1, 3]], 2),
    ("less", PEAK_X2, dict(axis=1), [[0, 2], [2, 1]], 2),
    ("greater", PEAK_X1, dict(), [[3], [6]], 2),
    ("greater", PEAK_X1, dict(order=3), [[3]], 1),
    ("greater", PEAK_X2, dict(), [[1, 0]], 1),
    ("greater", PEAK_X2, dict(axis=1), [[0, 1]], 1),
]
