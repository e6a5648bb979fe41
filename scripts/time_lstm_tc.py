import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stove_b200 import ops, _native as N
n, K, H, steps = 2048, 1024, 256, 3
dev = 'cuda'
x = torch.rand(n, K, device=dev)
ps = [(torch.rand(4 * H, K, device=dev) - 0.5) * 0.1, (torch.rand(4 * H, H, device=dev) - 0.5) * 0.1,
      torch.zeros(4 * H, device=dev), torch.zeros(4 * H, device=dev)]
lib = N.lib()
with torch.no_grad():
    for _ in range(3):
        ops.LstmEncoder.apply(x, *ps, steps)
    torch.cuda.synchronize()
    lib.stove_profile_enable(1); N.profile_read()
    for _ in range(10):
        ops.LstmEncoder.apply(x, *ps, steps)
    lib.stove_profile_enable(0)
recs = [t for name, t in N.profile_read() if name == 'lstm_gemm_cell_fwd']
t0 = recs[0::3]; t1 = recs[1::3]; t2 = recs[2::3]
print('debug=%s  step0 %.1f us  step1 %.1f us  step2 %.1f us (medians)' % (os.environ.get('STOVE_LSTM_TC_DEBUG'),
      1e3 * sorted(t0)[5], 1e3 * sorted(t1)[5], 1e3 * sorted(t2)[5]))
