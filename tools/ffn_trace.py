"""Event timeline of the fused FFN kernel (CTA (0,0)): per (tile, chunk) item, when the MMA warp waited / issued and
when the epilogue groups ran.  Cycles are clock64 of that SM, printed relative to the first traced event."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import timewarp_b200 as tw
from timewarp_b200 import _lib
from oracle import flow_oracle as fo
from timewarp_b200.peptides import tetrapeptide_2olx

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20, 28)
cls = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # 1 fused FFN, 2 mixing kernel, 3 token-major fused attention, 4 feature-major fused attention
pep = tetrapeptide_2olx()
m = tw.custom_transformer_nvp_constructor(tw.kernel_transformer_nvp_config("bf16x3"))
m.load_state_dict(fo.synth_state_dict(fo.OracleConfig(), 0))
m = m.cuda().eval()
g = torch.Generator().manual_seed(0)
x = (torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.01 * torch.randn(B, 65, 3, generator=g)).cuda()
y = x + 0.02 * torch.randn(B, 65, 3, generator=g).cuda()
xv, yv = torch.randn(B, 65, 3, generator=g).cuda(), torch.randn(B, 65, 3, generator=g).cuda()
at = torch.tensor(pep.atom_types)[None].repeat(B, 1).cuda()
mask = torch.zeros(B, 65, dtype=torch.bool).cuda()
e = torch.zeros(0, 2, dtype=torch.long).cuda()
kw = dict(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, adj_list=e, edge_batch_idx=e[:, 0], masked_elements=mask)
with torch.no_grad():
    m.log_likelihood(**kw)
    torch.cuda.synchronize()
    NR = 4 if cls == 4 else 3
    buf = torch.zeros(NR * 1024 * 2 + 8 + 2 * 160, dtype=torch.int64, device="cuda")
    lib = _lib.load()
    lib.tw_debug_set_trace(cls, buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.log_likelihood(**kw)
    e1.record()
    torch.cuda.synchronize()
    print(f"traced pass: {e0.elapsed_time(e1):.3f} ms")
    lib.tw_debug_set_trace(cls, None)
raw = buf.cpu().numpy()
NR = 4 if cls == 4 else 3
t = raw[:NR * 2048].reshape(NR, 1024, 2)
c0, g0, c1, g1 = (int(v) for v in raw[NR * 2048:NR * 2048 + 4])
if g1 > g0:
    print(f"MMA warp: {c1 - c0} cycles in {g1 - g0} ns -> SM clock {1e3 * (c1 - c0) / (g1 - g0):.0f} MHz")
if cls == 1:
    ce, ge, cx, gx = (int(v) for v in raw[NR * 2048 + 4:NR * 2048 + 8])
    if gx > ge:
        print(f"CTA (0,0): entry -> exit {cx - ce} cycles = {gx - ge} ns; entry -> first MMA-loop stamp {c0 - ce} cycles; last MMA-loop stamp -> exit {cx - c1} cycles")
    per = raw[NR * 2048 + 8:NR * 2048 + 8 + 2 * 148].reshape(148, 2)
    if per[:, 0].min() > 0:
        e, x = per[:, 0], per[:, 1]
        t0 = e.min()
        print(f"all 148 CTAs (global timer, ns after the first entry): entries {int(e.min() - t0)}..{int(e.max() - t0)}, exits {int(x.min() - t0)}..{int(x.max() - t0)}; "
              f"lifetimes {int((x - e).min())}..{int((x - e).max())} (median {int(np.median(x - e))}); kernel span {int(x.max() - t0)}; "
              f"exits by cluster rank-0 CTA index: first 20 pairs {[int(v - t0) for v in x[0:40:2]]}")
t0 = min(int(t[r, 0, 1]) for r in range(NR) if t[r, 0, 1] > 0)
if cls == 3:
    names = {0: {0: 'mma: head top', 1: 'mma: scores landed', 2: 'mma: MMA1 issued', 3: 'mma2: Wc kb0 landed', 4: 'mma2: h_full kb0', 5: 'mma2: Wc kb1 landed', 6: 'mma2: h_full kb1', 7: 'mma: sample top (item=sample)', 8: 'mma: xb_full (item=sample)'},
             1: {0: 'epi: wait dm', 1: 'epi: dm_full', 2: 'epi: h arrived', 3: 'epi: LN done', 4: 'epi: init_do done'},
             2: {0: 'conv: wait xs (item=sample)', 1: 'conv: xs_full', 2: 'conv: xb_free', 3: 'conv: done'}}
elif cls == 5:
    names = {0: {0: 'mma: item top', 1: 'mma: a_full / ready', 2: 'mma: G2 top', 3: 'mma: h_full', 4: 'mma: G2 issued'},
             1: {0: 'silu: wait d1', 1: 'silu: d1_full', 2: 'silu: done'},
             2: {0: 'io: tile top (item=tile)', 1: 'io: a_free', 2: 'io: gathered next', 3: 'io: y_full', 4: 'io: rows out'}}
elif cls == 4:
    names = {0: {0: 'P: head top', 1: 'P: issued'},
             1: {0: 'epi: pt_full', 1: 'epi: converted', 2: 'epi: drained previous group (item = group)'},
             2: {0: 'M: head top', 1: 'M: h_full', 2: 'M: scores landed', 3: 'M: issued'},
             3: {0: 'build: top (item = group)', 1: 'build: tiles free', 2: 'build: tiles written', 3: 'LN: finished (item = group)', 4: 'prefetch issued', 5: 'LN: rows_out seen', 6: 'LN: batch loads issued', 7: 'LN: pre-LN row 0 arrived', 8: 'LN: x row 0 arrived'}}
elif cls == 2:
    names = {0: {0: 'mma: sample top', 1: 'mma: hs_full', 2: 'mma: head top', 3: 'mma: scores landed', 4: 'mma: head issued'},
             1: {2: 'epi0: wait d_full', 3: 'epi0: d_full', 5: 'epi0: staging free', 4: 'epi0: staged', 7: 'epi0: sync2', 6: 'epi0: stores issued'},
             2: {0: 'conv: sample top (item=sample)', 2: 'conv: buffer free', 1: 'conv: done'}}
else:
  names = {0: {0: "mma: loop top", 1: "mma: W1hi landed", 2: "mma: G1 issued+committed", 3: "mma: W2 landed", 4: "mma: h_full[0] -> G2a", 5: "mma: h_full[1] -> G2b"},
           1: {0: "epi0: wait d1", 1: "epi0: d1_full", 2: "epi0: ld done", 3: "epi0: st done", 4: "epi0: arrived", 5: "epi0: LN got y_full (item=tile128)", 6: "epi0: LN stats done", 7: "epi0: LN y_free"},
           2: {0: "epi1: wait d1", 1: "epi1: d1_full", 2: "epi1: ld done", 3: "epi1: st done", 4: "epi1: arrived", 5: "epi1: got y_free", 6: "epi1: init_y done"}}
rows = []
for r in range(NR):
    for i in range(1024):
        code, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0:
            break
        ev, item = code & 0xFF, code >> 8
        if (cls == 5 and (lo <= (item if r < 2 else 2 * item) < hi)) or (cls != 5 and lo <= item < hi) or (cls == 4 and r == 3 and lo <= item * 6 + 5 < hi + 6) or (cls == 3 and ((r == 2) or (r == 0 and ev >= 7)) and lo <= item * 6 < hi):
            rows.append((clk - t0, item, names[r][ev]))
rows.sort()
prev = rows[0][0] if rows else 0
for clk, item, name in rows:
    print(f"{clk:9d} (+{clk - prev:5d})  item {item:3d}  {name}")
    prev = clk
# per-item period of the MMA warp
tops = [(int(t[0, i, 0]) >> 8, int(t[0, i, 1])) for i in range(1024) if t[0, i, 1] > 0 and (int(t[0, i, 0]) & 0xFF) == 0]
if len(tops) > 20:
    d = np.diff([c for _, c in tops])
    print("MMA-warp iteration period: median", np.median(d), "mean", d.mean(), "min", d.min(), "max", d.max(), "n", len(d))
