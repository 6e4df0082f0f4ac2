"""How fast / how accurate are fnet+cnet under reduced precision + channels_last? (not hot path; informs DESIGN)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import H, ITERS, W, _pairs, _state_dict
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model, _ = _state_dict()
model = model.to(dev).eval()
a, b = _pairs(1, dev)[0]
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
with torch.no_grad():
    torch.backends.cudnn.allow_tf32 = False
    ref = model._encoders(a, b)
    print("fp32 no-tf32: %.2f ms" % timeit(lambda: model._encoders(a, b)))
    torch.backends.cudnn.allow_tf32 = True
    out = model._encoders(a, b)
    print("tf32: %.2f ms  max|d| fmap %.4f cnet %.4f" % (timeit(lambda: model._encoders(a, b)), (out[0]-ref[0]).abs().max(), (out[2]-ref[2]).abs().max()))
    for dt in (torch.bfloat16, torch.float16):
        for cl in (False, True):
            m2 = model
            def enc():
                x1 = (2 * (a / 255.0) - 1.0); x2 = (2 * (b / 255.0) - 1.0)
                if cl:
                    x1 = x1.contiguous(memory_format=torch.channels_last); x2 = x2.contiguous(memory_format=torch.channels_last)
                with torch.autocast("cuda", dtype=dt):
                    f1, f2 = m2.fnet([x1, x2]); c = m2.cnet(x1)
                return f1.float(), f2.float(), c.float()
            if cl:
                m2.fnet.to(memory_format=torch.channels_last); m2.cnet.to(memory_format=torch.channels_last)
            out = enc()
            print("%s channels_last=%s: %.2f ms  max|d| fmap %.4f (ref absmax %.2f) cnet %.4f" % (dt, cl, timeit(enc), (out[0]-ref[0]).abs().max(), ref[0].abs().max(), (out[2]-ref[2]).abs().max()))
