import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphgps_b200 import _lib
from graphgps_b200.batch import batch_from_lists
from graphgps_b200.graph import graph_of
lib = _lib.load(); DEV = "cuda:0"
st = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
H, hd, n = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 40
hp = (hd + 15)//16*16
b = batch_from_lists([n], [[]], d=8).to(DEV); gs = graph_of(b); N = n; D = H*hd
QKV = torch.randn(N, 3*D, device=DEV)
x = torch.zeros(N, 3*H, hp, device=DEV); x[:, :, :hd] = QKV.view(N, 3*H, hd); x = x.view(N, 3*H*hp)
hi = x.to(torch.bfloat16); lo = (x - hi.float()).to(torch.bfloat16); pl = torch.stack([hi, lo]).contiguous()
dbg = torch.zeros(3, 128, 128, device=DEV)
lib.gps_debug_attn(dbg.data_ptr())
O = torch.full((N, D), float("nan"), device=DEV); lse = torch.empty(N, H, device=DEV)
rc = lib.gps_attention_forward_tc(C.byref(gs.desc), H, hd, pl[0].data_ptr(), pl[1].data_ptr(), 3*H*hp, O.data_ptr(), D, lse.data_ptr(), 0.0, 0, 0, 0, st())
torch.cuda.synchronize(); lib.gps_debug_attn(0)
Q, K, V = QKV[:, :D].double(), QKV[:, D:2*D].double(), QKV[:, 2*D:].double()
S = (Q @ K.t()).cpu()
Sg = dbg[0, :n, :n].double().cpu(); Pg = dbg[1, :n, :n].double().cpu(); Og = dbg[2, :n, :hd].double().cpu()
print("S err", float((Sg - S).abs().max()), " S[0,:4] got", Sg[0, :4].tolist(), "ref", S[0, :4].tolist())
Pref = torch.exp(S / hd**0.5 - (S / hd**0.5).max(1, keepdim=True).values)
print("P err", float((Pg - Pref).abs().max()))
Oexp = Pg @ V.cpu()
print("O(acc) vs P_dumped@V err", float((Og - Oexp).abs().max()), "O acc[0,:4]", Og[0, :4].tolist(), "exp", Oexp[0, :4].tolist())
# which keys contribute? solve per-key weights by least squares on row 0
for mask_name, idx in (("keys j%16<8", [j for j in range(n) if j % 16 < 8]), ("keys j%16>=8", [j for j in range(n) if j % 16 >= 8]), ("j<32", list(range(min(32, n))))):
    e = (Og - Pg[:, idx] @ V.cpu()[idx]).abs().max()
    print("  subset", mask_name, "err", float(e))
os.makedirs("gpurun_out", exist_ok=True)
torch.save({"dbg": dbg.cpu(), "QKV": QKV.cpu(), "O": O.cpu(), "hd": hd, "n": n}, f"gpurun_out/attn_dump_{hd}_{n}.pt")
