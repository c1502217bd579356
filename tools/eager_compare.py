"""The 'existing Blackwell kernels' comparator: the same UtNet layers run by PyTorch eager
(cuDNN / cuBLAS) on the same B200 — fp32 NCHW as the reference script would run it on a GPU, and bf16
channels_last — next to this library's forward, on batches of crops.  Measurement only; not on any product
path.   python tools/eager_compare.py [cs] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb  # noqa: E402

cs = int(sys.argv[1]) if len(sys.argv) > 1 else 248
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 28
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = nb.UtNet().to(dev).eval()       # holds ordinary nn.Conv2d / ConvTranspose2d / PReLU children


def eager_forward(m, x):
    """What networks/UtNet.py:97-109 computes, with the module's own torch layers."""
    cat = torch.cat
    x = m.pad(x)
    l1 = m.convs1(x)
    l2 = m.convs2(m.maxpool(l1))
    l3 = m.convs3(m.maxpool(l2))
    l4 = m.convs4(m.maxpool(l3))
    y = cat([m.up1(m.bottom(m.maxpool(l4))), l4], dim=1)
    y = cat([m.up2(m.tconvs1(y)), l3], dim=1)
    y = cat([m.up3(m.tconvs2(y)), l2], dim=1)
    y = cat([m.up4(m.tconvs3(y)), l1], dim=1)
    return m.unpad(m.tconvs4(y))


def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


x = torch.rand(batch, 3, cs, cs, device=dev)
ucs = cs - 24
mp_per_crop = (ucs - 6) ** 2 / 1e6      # image pixels a crop contributes at overlap 6
rows = []
with torch.no_grad():
    ref = eager_forward(model, x)
    ours = model(x)
    print(f"cs {cs} batch {batch}: max |ours - eager fp32| = {float((ours - ref).abs().max()):.3e} "
          f"(sigma_out {float(ref.std()):.4f})")
    ms = timed(lambda: model(x), 10)
    rows.append(("nind_denoise_b200 (bf16, tcgen05)", ms))
    torch.backends.cudnn.benchmark = True
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ms = timed(lambda: eager_forward(model, x), 3)
        rows.append((f"torch eager fp32 NCHW (cuDNN, tf32={'on' if tf32 else 'off'})", ms))
    import copy
    m16 = copy.deepcopy(model).to(torch.bfloat16).to(memory_format=torch.channels_last)
    x16 = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ms = timed(lambda: eager_forward(m16, x16), 5)
    rows.append(("torch eager bf16 channels_last (cuDNN)", ms))
for name, ms in rows:
    print(f"  {name:52s} {ms:9.3f} ms/batch  {ms / batch * 1e3:8.1f} us/crop  ~{batch * mp_per_crop / ms * 1e3:7.1f} MP/s of image")
