import sys; sys.path.insert(0, ".")
import gpirt_b200.sampler as G
print("int8 peak TOP/s:", G.int8_peak_tops(), "fp64 peaks:", G.fp64_peak_tflops())
