import time, torch, sys, os, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientvlm_b200 import _lib, kernels as K
dev = torch.device("cuda", 0)
def bench(name, fn, n=3000):
    for _ in range(100): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("%-50s %.2f us" % (name, (t1 - t0) / n * 1e6))
bench("torch.empty small bf16", lambda: torch.empty((320, 768), device=dev, dtype=torch.bfloat16))
bench("torch.empty 20MB", lambda: torch.empty((25216, 768), device=dev, dtype=torch.bfloat16))
bench("torch.empty 600MB", lambda: torch.empty((128*12*197*197,), device=dev, dtype=torch.float32))
keep = []
def alloc_keep():
    keep.append(torch.empty((25216, 768), device=dev, dtype=torch.bfloat16))
    if len(keep) > 200: keep.clear()
bench("torch.empty 38MB keep200", alloc_keep)
bench("torch.empty 'cuda' str", lambda: torch.empty((320, 768), device="cuda", dtype=torch.bfloat16))
bench("torch.zeros small", lambda: torch.zeros((320,), device=dev))
bench("current_stream().cuda_stream", lambda: torch.cuda.current_stream().cuda_stream)
bench("_cuda_getCurrentRawStream", lambda: torch._C._cuda_getCurrentRawStream(0))
lib = _lib.load()
bench("ctypes evlm_launch_count", lambda: lib.evlm_launch_count())
x = torch.randn(320, 768, device=dev); y = torch.empty(320, 768, device=dev, dtype=torch.bfloat16)
bench("K.cast_bf16 320x768", lambda: K.cast_bf16(x, y) if False else K.cast_bf16(x))
class Noop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a): return a.view_as(a)
    @staticmethod
    def backward(ctx, g): return g
xr = x.clone().requires_grad_(True)
bench("autograd Function apply (noop)", lambda: Noop.apply(xr))
print("gc thresholds", gc.get_threshold(), "alloc conf", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"))
