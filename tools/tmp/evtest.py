import torch, inspect
print(torch.__version__)
print(inspect.signature(torch.cuda.Event.__new__) if hasattr(torch.cuda.Event,'__new__') else '')
try:
    a=torch.cuda.Event(enable_timing=True, external=True); b=torch.cuda.Event(enable_timing=True, external=True)
except Exception as ex:
    print("no external:", ex); raise SystemExit
x=torch.zeros(1<<26, device='cuda')
s=torch.cuda.Stream()
g=torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    x.add_(1)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        a.record()
        x.add_(1); x.mul_(2)
        b.record()
for i in range(3):
    g.replay()
    torch.cuda.synchronize()
    try:
        print("elapsed", a.elapsed_time(b))
    except Exception as ex:
        print("elapsed failed:", ex)
