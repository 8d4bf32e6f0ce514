import torch, time
d = torch.empty(65536*1000, dtype=torch.float64, device="cuda")
h = torch.empty(65536*1000, dtype=torch.float64).pin_memory()
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); h.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); print(f"D2H 524 MB pinned: {ms:.2f} ms = {d.numel()*8/ms/1e6:.1f} GB/s")
e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); print(f"H2D 524 MB pinned: {ms:.2f} ms = {d.numel()*8/ms/1e6:.1f} GB/s")
