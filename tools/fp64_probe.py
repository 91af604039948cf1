"""FP64-pipe ceiling under realistic operand patterns and occupancies (experiment; run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microlensing_b200 as mb
ctx = mb._capi.context(0)
nominal = 148 * 64 * 1.965e9
print("DFMA-chain peak (constant operands): %.2f TFLOP/s = %.3f of nominal" % (ctx.fp64_peak_tflops(), ctx.fp64_peak_tflops() / 37.2))
for mode in (1, 2):
    for bps in (2, 3, 4, 6, 8, 12, 16):
        r = ctx.fp64_probe(mode, bps)
        print(f"mode {mode} blocks/SM {bps:2d} ({bps * 4} warps/SM): {r:.3e} inst/s = {r / nominal:.3f} of nominal issue rate")
