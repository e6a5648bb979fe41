"""In-graph GPU time of the encoder GEMMs: 3 accumulated TF32 calls vs one K-concatenated call."""
import torch
dev = 'cuda'
torch.backends.cuda.matmul.allow_tf32 = True

def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * reps) * 1000

n, K, H = 2048, 1024, 256
cases = {
    'gx    (n,K)@(K,4H)':    ((n, K), (K, 4 * H), False),
    'gh    (n,H)@(H,4H)':    ((n, H), (H, 4 * H), False),
    'dh    (n,4H)@(4H,H)':   ((n, 4 * H), (4 * H, H), False),
    'dWhh  (4H,n)@(n,H)':    ((n, 4 * H), (n, H), True),
    'dWih  (4H,n)@(n,K)':    ((n, 4 * H), (n, K), True),
}
for name, (sa, sb, ta) in cases.items():
    A = torch.randn(*sa, device=dev); B = torch.randn(*sb, device=dev)
    Am = A.t() if ta else A
    M, Kc = Am.shape; N = B.shape[1]
    C = torch.empty(M, N, device=dev)
    def three():
        torch.mm(Am, B, out=C); C.addmm_(Am, B); C.addmm_(Am, B)
    # K-concatenated operands
    if ta:
        A3 = torch.randn(3 * sa[0], sa[1], device=dev).t(); B3 = torch.randn(3 * sb[0], sb[1], device=dev)
    else:
        A3 = torch.randn(sa[0], 3 * sa[1], device=dev); B3 = torch.randn(3 * sb[0], sb[1], device=dev)
    def one():
        torch.mm(A3, B3, out=C)
    torch.backends.cuda.matmul.allow_tf32 = False
    t32 = graph_time(lambda: torch.mm(Am, B, out=C))
    torch.backends.cuda.matmul.allow_tf32 = True
    print('%-24s fp32 %.1f us | 3 x tf32 %.1f us | 1 x tf32 (3K) %.1f us' % (name, t32, graph_time(three), graph_time(one)))
