"""Launch only pack + the K1 lag kernel on the BASELINE config-2 shape (for ncu captures).  Scratch tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import _lib  # noqa: E402

lib = _lib.load()
nC, nF = 5, 200000
nR = int(os.environ.get("CT_NR", "76"))
g = torch.Generator(device="cuda").manual_seed(1)
vt = torch.randn((nC, nF, nR, 3), device="cuda", generator=g)
vt = (vt / vt.norm(dim=-1, keepdim=True)).contiguous()
L = nF // 2
pitch = lib.sr_ct_row_pitch(nF)
packed = torch.empty((nR, nC, 3, pitch), dtype=torch.float32, device="cuda")
S = torch.empty((nR, nC, L), dtype=torch.float64, device="cuda")
st = _lib.current_stream_ptr()
_lib.check(lib.sr_pack_vectors_f32(vt.data_ptr(), nC, nF, nR, None, packed.data_ptr(), pitch, st))
for _ in range(int(os.environ.get("CT_REPS", "2"))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.sr_ct_lag_sums(packed.data_ptr(), pitch, nC, nF, nR, L, S.data_ptr(), st))
    b.record()
    torch.cuda.synchronize()
    print("ct_lag ms", a.elapsed_time(b))
