import torch, time
a = torch.randn(8192, 8192, device='cuda'); 
d = torch.empty(100*1024*1024, dtype=torch.float64, device='cuda'); h = torch.empty_like(d, device='cpu').pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def compute():
    with torch.cuda.stream(s1):
        for _ in range(40): b = a @ a
def copy():
    with torch.cuda.stream(s2):
        h.copy_(d, non_blocking=True)
for f in (compute, copy): f(); torch.cuda.synchronize()
for name, fs in (("compute", [compute]), ("copy", [copy]), ("both", [copy, compute])):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for f in fs: f()
    torch.cuda.synchronize(); print(name, round(time.perf_counter() - t0, 4))
