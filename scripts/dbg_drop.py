import sys, torch as th
sys.path.insert(0, '/root/repo')
from dreamgnn_b200 import ops as o
dev = th.device('cuda:0')
e, n_d, n_s, p = 20000, 50, 40, 0.3
pairs = o.PairGraph(th.randint(0, n_d, (e,), device=dev), th.randint(0, n_s, (e,), device=dev), n_d, n_s)
lib = o.L.load()
big = th.full((n_d, 128), 5.0, device=dev); zero = th.zeros(n_s, 128, device=dev)
b2 = th.zeros(64, device=dev); w3 = th.ones(64, device=dev); b3 = th.zeros(1, device=dev)
rates = []
for k in range(128):
    w2 = th.zeros(64, 128, device=dev)
    w2[:, k] = 1.0                      # every output unit j sees z1[k]
    z2 = th.empty(e, 64, device=dev); outp = th.empty(e, device=dev)
    o.L.check(lib.dg_decoder_fwd_f32(o.L.ptr(pairs.src), o.L.ptr(pairs.dst), e, o.L.ptr(big), o.L.ptr(zero), o.L.ptr(w2),
              o.L.ptr(b2), o.L.ptr(w3), o.L.ptr(b3), p, 42, o.L.ptr(outp), o.L.ptr(z2), o.L.stream()), 'fwd')
    any_kept = (z2 > 0).any(1).float().mean()      # P(keep1[k]) * P(any of 64 keep2) ~= P(keep1[k])
    rates.append(round(float(any_kept), 3))
print('keep1 per unit:', rates[:16], '...', rates[60:70])
print('keep2 per unit:', [round(float(x), 3) for x in (z2 > 0).float().sum(0)[:8] / (z2 > 0).any(1).float().sum()])
for k in (0, 1, 2, 5, 64, 65):
    w2 = th.zeros(64, 128, device=dev); w2[:, k] = 1.0
    z2 = th.empty(e, 64, device=dev); outp = th.empty(e, device=dev)
    o.L.check(lib.dg_decoder_fwd_f32(o.L.ptr(pairs.src), o.L.ptr(pairs.dst), e, o.L.ptr(big), o.L.ptr(zero), o.L.ptr(w2),
              o.L.ptr(b2), o.L.ptr(w3), o.L.ptr(b3), p, 42, o.L.ptr(outp), o.L.ptr(z2), o.L.stream()), 'fwd')
    print('k=%d: P(z2[:,j]>0) j=0..9:' % k, [round(float(x), 3) for x in (z2 > 0).float().mean(0)[:10]])
w2 = th.ones(64, 128, device=dev)
z2 = th.empty(e, 64, device=dev); outp = th.empty(e, device=dev)
o.L.check(lib.dg_decoder_fwd_f32(o.L.ptr(pairs.src), o.L.ptr(pairs.dst), e, o.L.ptr(big), o.L.ptr(zero), o.L.ptr(w2),
          o.L.ptr(b2), o.L.ptr(w3), o.L.ptr(b3), p, 42, o.L.ptr(outp), o.L.ptr(z2), o.L.stream()), 'fwd')
s = 65536.0 / (65536.0 - round(0.3 * 65536))
nk = z2.max(1).values / (5 * s * s)
print('ones: kept2', float((z2 > 0).float().mean()), 'mean z2 kept', float(z2[z2 > 0].mean()), 'expect', 128 * 0.7 * 5 * s * s,
      'nkept mean/min/max', float(nk.mean()), float(nk.min()), float(nk.max()))
