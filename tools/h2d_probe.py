import torch, time
n = 33_300_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
src = torch.randint(0, 255, (n,), dtype=torch.uint8)
def t_copy():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for i in range(3): t_copy()
print("cold (not rewritten):", [round(t_copy(), 3) for _ in range(3)])
torch.set_num_threads(16)
for k in range(3):
    h.copy_(src)   # CPU write (multi-threaded memcpy)
    print("right after a CPU rewrite:", round(t_copy(), 3), "then", round(t_copy(), 3))
h16 = torch.empty(n // 2, dtype=torch.int16).pin_memory(); d16 = torch.empty(n // 2, dtype=torch.int16, device='cuda')
h16.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); d16.copy_(h16, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("zeros int16:", round(e0.elapsed_time(e1), 3))
