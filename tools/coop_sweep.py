import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlx8_ws_audio_transformer_b200 import LogMelFrontend
from mlx8_ws_audio_transformer_b200 import _native as N
from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank
def front(nm, pairs, variant=3):
    if pairs: os.environ["LM_TF_PAIRS_PER_CLIP"] = str(pairs)
    else: os.environ.pop("LM_TF_PAIRS_PER_CLIP", None)
    return LogMelFrontend(400, 160, slaney_mel_filter_bank(201, nm), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True, variant=variant)
def timeit(fe, x, out, iters=8):
    for _ in range(3): fe.forward(x, out=out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fe.forward(x, out=out); ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(4096, 480000, generator=g, device="cuda") * 0.1
out = torch.empty(4096, nm, 3000, device="cuda")
f1, f4, fa = front(nm, 1), front(nm, 4), front(nm, 0)
os.environ["LM_TF_MIN_BATCH"] = "1000000000"; ft = front(nm, 0, 0); os.environ.pop("LM_TF_MIN_BATCH")
# correctness: both modes bit-identical
a = f1.forward(x[:700]).clone(); b = f4.forward(x[:700])
print("pair vs cta mode bit-identical:", bool(torch.equal(a, b)))
for B in (37, 74, 148, 222, 296, 444, 592, 600, 740, 888, 1024, 1184, 1480, 2048, 4096):
    t1, t4, ta, tt = (timeit(f, x[:B], out[:B]) for f in (f1, f4, fa, ft))
    print(f"batch {B:5d}: clip/pair {t1*1e3:8.1f} us  clip/CTA {t4*1e3:8.1f} us  auto {ta*1e3:8.1f} us  CTA-tiled kernel {tt*1e3:8.1f} us")
