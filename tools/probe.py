import sys
sys.path.insert(0, '.')
from crescent_credentials_b200 import ffi
ctx = ffi.Context(0)
print('probes G/s: imad32 %.0f wide %.0f frmul %.2f fqmul %.2f' % tuple(ctx.bench_int_pipe(i) for i in range(4)))
