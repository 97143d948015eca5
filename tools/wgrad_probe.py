import torch, time
torch.backends.cuda.matmul.allow_tf32 = True
n = 65536
for k, c in [(256, 256), (54, 256), (256, 21)]:
    x = torch.randn(n, k, device="cuda"); dz = torch.randn(n, c, device="cuda")
    ref = (x.double().t() @ dz.double())
    variants = {
        "x.t() @ dz": lambda: x.t() @ dz,
        "(dz.t() @ x).t()": lambda: (dz.t() @ x).t(),
        "x.t().contiguous() @ dz": lambda: x.t().contiguous() @ dz,
        "bf16 x.t() @ dz": lambda: (x.bfloat16().t() @ dz.bfloat16()).float(),
        "einsum": lambda: torch.einsum("nk,nc->kc", x, dz),
        "splitK bmm 16": lambda: torch.bmm(x.view(16, n // 16, k).transpose(1, 2), dz.view(16, n // 16, c)).sum(0),
        "splitK bmm 64": lambda: torch.bmm(x.view(64, n // 64, k).transpose(1, 2), dz.view(64, n // 64, c)).sum(0),
    }
    for name, f in variants.items():
        out = f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        print(f"k={k:3d} c={c:3d} {name:28s} {e0.elapsed_time(e1)/20*1e3:8.1f} us  rel err {err:.1e}")
