"""Kernel-level parity on the GPU: every C-ABI entry point against a plain PyTorch fp32 (or the
float64 oracle) restatement of the same op, on seeded inputs.  fp32 tier tolerance: 1e-3
relative (north star); most kernels land at 1e-6.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import _golden as g

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
  from vision4leg_b200 import engine
  return engine, engine.ops_for(DEV)


def rel(a, b):
  a, b = a.double().cpu(), b.double().cpu()
  return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def test_conv_stack_forward_backward_matches_torch():
  engine, ops = _ops()
  torch.manual_seed(0)
  B, S = 5, 7
  x = torch.randn(B, S + 16384, device=DEV)
  idx = torch.tensor([3, 0, 4, 1, 2], device=DEV, dtype=torch.int32)
  conv = torch.nn.Sequential(torch.nn.Conv2d(4, 32, 8, 4), torch.nn.ReLU(), torch.nn.Conv2d(32, 64, 4, 2),
                             torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, 1), torch.nn.ReLU()).to(DEV)
  P = {"t.layers.%d.%s" % (i, n): getattr(conv[i], n).data for i in (0, 2, 4) for n in ("weight", "bias")}
  plan = engine._Plan(ops, 1)
  trunk = engine._ConvTrunk(plan, "t.")
  inp = engine.Input(B, x, S + 16384, 0, x, S + 16384, S, idx)
  a3 = trunk.forward(P, inp)
  img = x[idx.long(), S:].reshape(B, 4, 64, 64).clone().requires_grad_(True)
  ref = conv(img)
  assert rel(a3.reshape(B, 4, 4, 64).permute(0, 3, 1, 2), ref) < 1e-5
  assert rel(trunk.a1.reshape(B, 15, 15, 32).permute(0, 3, 1, 2), conv[1](conv[0](img))) < 1e-5
  # backward: random upstream gradient on the post-ReLU output
  gout = torch.randn_like(ref)
  ref.backward(gout)
  da3 = (gout * (ref > 0)).permute(0, 2, 3, 1).reshape(B, 16, 64).contiguous()
  G = {k: torch.full_like(v, float("nan")) for k, v in P.items()}
  trunk.backward(P, G, da3)
  for i in (0, 2, 4):
    assert rel(G["t.layers.%d.weight" % i], conv[i].weight.grad) < 2e-5, i
    assert rel(G["t.layers.%d.bias" % i], conv[i].bias.grad) < 2e-5, i


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 12, 93), (130, 256, 1024), (1000, 70, 33)])
def test_linear_fwd_dgrad_wgrad(M, N, K):
  engine, ops = _ops()
  RM = engine.RM
  torch.manual_seed(M + N + K)
  x = torch.randn(M, K, device=DEV)
  w = torch.randn(N, K, device=DEV) / math.sqrt(K)
  b = torch.randn(N, device=DEV)
  y = torch.empty(M, N, device=DEV)
  ops.linear_fwd(x, RM.dense(K), None, w, b, y, RM.dense(N), M, N, K, True)
  ref = F.relu(F.linear(x, w, b))
  assert rel(y, ref) < 1e-5
  dy = torch.randn(M, N, device=DEV)
  dx = torch.empty(M, K, device=DEV)
  ops.linear_dgrad(dy, RM.dense(N), w, dx, RM.dense(K), M, N, K, mask=x, mask_map=RM.dense(K))
  assert rel(dx, (dy @ w) * (x > 0)) < 1e-5
  dx2 = torch.ones(M, K, device=DEV)
  ops.linear_dgrad(dy, RM.dense(N), w, dx2, RM.dense(K), M, N, K, accum=True)
  assert rel(dx2, dy @ w + 1) < 1e-5
  dw = torch.empty(N, K, device=DEV)
  db = torch.empty(N, device=DEV)
  ops.linear_wgrad(dy, RM.dense(N), x, RM.dense(K), None, dw, db, M, N, K)
  assert rel(dw, dy.t() @ x) < 1e-5
  assert rel(db, dy.sum(0)) < 1e-5


def test_wgrad_is_deterministic_and_splits_large_m():
  engine, ops = _ops()
  RM = engine.RM
  torch.manual_seed(1)
  M, N, K = 50000, 32, 256
  x = torch.randn(M, K, device=DEV)
  dy = torch.randn(M, N, device=DEV)
  outs = []
  for _ in range(2):
    dw = torch.empty(N, K, device=DEV); db = torch.empty(N, device=DEV)
    ops.linear_wgrad(dy, RM.dense(N), x, RM.dense(K), None, dw, db, M, N, K)
    outs.append((dw.clone(), db.clone()))
  assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
  assert rel(outs[0][0], (dy.double().t() @ x.double())) < 1e-5


@pytest.mark.parametrize("B,T,d,nh", [(3, 17, 64, 1), (2, 16, 64, 4), (1, 5, 32, 2)])
def test_attention_fwd_bwd(B, T, d, nh):
  engine, ops = _ops()
  torch.manual_seed(B * T)
  qkv = torch.randn(B, T, 3 * d, device=DEV, requires_grad=True)
  o = torch.empty(B, T, d, device=DEV)
  p = torch.empty(B, nh, T, T, device=DEV)
  ops.attn_fwd(qkv.detach(), o, p, B, T, d, nh)
  q, k, v = qkv.split(d, -1)
  hd = d // nh
  sp = lambda t: t.reshape(B, T, nh, hd).transpose(1, 2)
  pr = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / math.sqrt(hd), -1)
  ref = (pr @ sp(v)).transpose(1, 2).reshape(B, T, d)
  assert rel(o, ref) < 1e-5 and rel(p, pr) < 1e-5
  go = torch.randn_like(ref)
  ref.backward(go)
  dqkv = torch.empty_like(qkv)
  ops.attn_bwd(qkv.detach(), p, go, dqkv, B, T, d, nh)
  assert rel(dqkv, qkv.grad) < 2e-5


@pytest.mark.parametrize("rows,d", [(1, 64), (51, 64), (1000, 256), (9, 33)])
def test_layernorm_fwd_bwd(rows, d):
  engine, ops = _ops()
  torch.manual_seed(rows)
  a = torch.randn(rows, d, device=DEV, requires_grad=True)
  r = torch.randn(rows, d, device=DEV, requires_grad=True)
  gm = torch.randn(d, device=DEV, requires_grad=True)
  bt = torch.randn(d, device=DEV, requires_grad=True)
  y = torch.empty(rows, d, device=DEV); z = torch.empty(rows, d, device=DEV)
  st = torch.empty(rows, 2, device=DEV)
  ops.ln_fwd(a.detach(), r.detach(), gm.detach(), bt.detach(), y, z, st, rows, d)
  ref = F.layer_norm(a + r, (d,), gm, bt, 1e-5)
  assert rel(y, ref) < 1e-5
  gy = torch.randn_like(ref)
  ref.backward(gy)
  dz = torch.empty(rows, d, device=DEV); dg = torch.empty(d, device=DEV); db = torch.empty(d, device=DEV)
  ops.ln_bwd(gy, z, st, gm.detach(), dz, dg, db, rows, d)
  assert rel(dz, a.grad) < 2e-5 and rel(dg, gm.grad) < 2e-5 and rel(db, bt.grad) < 2e-5


def test_pool_fwd_bwd():
  engine, ops = _ops()
  torch.manual_seed(2)
  B, T, d = 6, 17, 64
  tok = torch.randn(B, T, d, device=DEV, requires_grad=True)
  for mode, ref in ((0, torch.cat([tok[:, 0], tok[:, 1:].mean(1)], -1)), (1, tok.mean(1))):
    out = torch.empty_like(ref)
    ops.pool_fwd(tok.detach(), out, B, T, d, mode)
    assert rel(out, ref) < 1e-6
    gout = torch.randn_like(ref)
    tok.grad = None
    ref.backward(gout, retain_graph=True)
    dtok = torch.empty(B, T, d, device=DEV)
    ops.pool_bwd(gout, dtok, B, T, d, mode)
    assert rel(dtok, tok.grad) < 1e-6


# -------------------------------------------------------------------------------------------------
# GAE
# -------------------------------------------------------------------------------------------------
def _gae_gpu(roll, last_value, tlf, mode=0):
  engine, ops = _ops()
  T, E = roll["rewards"].shape[:2]
  up = lambda k: torch.tensor(np.ascontiguousarray(roll[k], np.float32).reshape(T, -1), device=DEV)
  r, v, d, tl = up("rewards"), up("values"), up("terminals"), up("time_limits")
  lv = torch.tensor(np.asarray(last_value, np.float32).reshape(E), device=DEV)
  advs = torch.empty(T, E, device=DEV); rets = torch.empty(T, E, device=DEV)
  st, se = (E, 1) if tl.shape[1] == E and E > 1 else (1, 0)
  ops.gae(r, v, d, tl, st, se, lv, advs, rets, T, E, 0.99, 0.95 if mode == 0 else 1.0, tlf, mode)
  return advs.cpu().numpy().reshape(T, E, 1), rets.cpu().numpy().reshape(T, E, 1)


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_gae_matches_reference_golden(case):
  G = g.load("gae")
  roll, last_value, tlf = g.gae_case(case, G["gae_%s/cfg" % case])
  # the kernel takes fp32 inputs: compare against the golden float64 result computed by the
  # reference from the same fp32-representable values (last_value is rounded to fp32 here)
  advs, rets = _gae_gpu(roll, last_value, tlf, 0)
  assert g.rel_err(advs, G["gae_%s/advs" % case]) < 1e-6
  assert g.rel_err(rets, G["gae_%s/rets" % case]) < 1e-6
  advs, rets = _gae_gpu(roll, last_value, tlf, 1)
  assert g.rel_err(advs, G["disc_%s/advs" % case]) < 1e-6
  assert g.rel_err(rets, G["disc_%s/rets" % case]) < 1e-6


def test_gae_full_size_matches_oracle_and_is_segmented():
  """1M-transition sweep shape (T=131072, E=8): multi-chunk path vs the float64 oracle, plus the
  size-independent property that an episode end cuts the recurrence (advantages before a
  terminal do not depend on anything after it)."""
  from oracle import ppo_oracle as po, synth
  T, E = 131072, 8
  roll = synth.make_rollout(9, T, E, 1, 1, with_img=False, p_term=1 / 500.0, time_limit_p=0.001)
  lv = np.random.default_rng(1).standard_normal((E, 1)).astype(np.float32)
  advs, rets = _gae_gpu(roll, lv, True, 0)
  ra, rr = po.gae(roll["rewards"], roll["values"], roll["terminals"], roll["time_limits"], lv, 0.99, 0.95, True)
  assert g.rel_err(advs, ra) < 1e-6 and g.rel_err(rets, rr) < 1e-6
  # segmentation property
  t_cut = int(np.argmax(roll["terminals"][:, 0, 0] > 0))
  roll2 = {k: np.array(v, copy=True) for k, v in roll.items()}
  roll2["rewards"][t_cut + 1:, 0] += 100.0
  advs2, _ = _gae_gpu(roll2, lv, True, 0)
  np.testing.assert_array_equal(advs2[:t_cut + 1, 0], advs[:t_cut + 1, 0])
  assert not np.array_equal(advs2[t_cut + 1:, 0], advs[t_cut + 1:, 0])
  np.testing.assert_array_equal(advs2[:, 1:], advs[:, 1:])      # other env columns untouched


def test_gae_linearity_in_rewards():
  """A(r1 + r2, V=0) = A(r1, V=0) + A(r2, V=0): size-independent property of the scan."""
  from oracle import synth
  T, E = 4096, 4
  a = synth.make_rollout(1, T, E, 1, 1, with_img=False, p_term=0.01)
  b = synth.make_rollout(2, T, E, 1, 1, with_img=False, p_term=0.01)
  for r in (a, b):
    r["values"][:] = 0
    r["terminals"] = a["terminals"]
  lv = np.zeros((E, 1), np.float32)
  s = {k: np.array(v, copy=True) for k, v in a.items()}
  s["rewards"] = a["rewards"] + b["rewards"]
  A1, _ = _gae_gpu(a, lv, False)
  A2, _ = _gae_gpu(b, lv, False)
  A3, _ = _gae_gpu(s, lv, False)
  assert g.rel_err(A3, A1 + A2) < 1e-5


# -------------------------------------------------------------------------------------------------
# losses and optimiser
# -------------------------------------------------------------------------------------------------
def test_pf_vf_loss_and_adam_match_torch():
  from vision4leg_b200._lib import INFO_KEYS, INFO_STRIDE
  engine, ops = _ops()
  torch.manual_seed(5)
  n, A, N = 300, 12, 1000
  idx = torch.randperm(N, device=DEV)[:n].to(torch.int32)
  il = idx.long()
  acts = 0.2 * torch.randn(N, A, device=DEV)
  adv = torch.randn(N, device=DEV)
  rets = torch.randn(N, device=DEV)
  vold = torch.randn(N, device=DEV)
  mean = (0.05 * torch.randn(n, A, device=DEV)).requires_grad_(True)
  tmean = mean.detach() + 0.01 * torch.randn(n, A, device=DEV)
  logstd = (math.log(0.125) + 0.1 * torch.randn(A, device=DEV))
  logstd[0] = 2.5       # outside the clamp: gradient must be blocked
  logstd.requires_grad_(True)
  tlogstd = logstd.detach() + 0.01
  info = torch.zeros(2, INFO_STRIDE, device=DEV)
  slot = torch.ones(1, device=DEV, dtype=torch.int32)
  stats = torch.zeros(8, device=DEV, dtype=torch.float64)
  ops.adv_stats(adv, idx, n, stats)
  d_mean = torch.empty(n, A, device=DEV); d_ls = torch.empty(A, device=DEV)
  ops.pf_loss(mean.detach(), logstd.detach(), tmean, tlogstd, acts, adv, idx, stats, d_mean, d_ls, n, A,
              1.0 / n, 1.0 / n, 0.2, 0.005, info, slot)
  # torch restatement (reference ppo.py:42-92)
  from torch.distributions import Normal
  a = adv[il].unsqueeze(1)
  ah = (a - a.mean()) / (a.std() + 1e-5)
  ls = torch.clamp(logstd, -5, 2)
  dist = Normal(mean, torch.exp(ls).unsqueeze(0).expand_as(mean))
  lp = dist.log_prob(acts[il]).sum(-1, keepdim=True)
  ent = dist.entropy().sum(-1, keepdim=True)
  tlp = Normal(tmean, torch.exp(torch.clamp(tlogstd, -5, 2)).expand_as(tmean)).log_prob(acts[il]).sum(-1, keepdim=True)
  ratio = torch.exp(lp - tlp)
  loss = -torch.min(torch.clamp(ratio, 0.8, 1.2) * ah, ratio * ah).mean() - 0.005 * ent.mean()
  loss.backward()
  assert rel(d_mean, mean.grad) < 1e-4
  assert rel(d_ls, logstd.grad) < 1e-4 and float(d_ls[0]) == 0.0
  row = dict(zip(INFO_KEYS, info[1].tolist()))
  exp = {"advs/mean": a.mean(), "advs/std": a.std(), "advs/max": a.max(), "advs/min": a.min(),
         "Training/policy_loss": loss, "logprob/mean": lp.mean(), "logprob/std": lp.std(),
         "logprob/max": lp.max(), "logprob/min": lp.min(), "log_std/mean": ls.mean(),
         "log_std/std": ls.std(), "log_std/max": ls.max(), "log_std/min": ls.min(),
         "ratio/max": ratio.max(), "ratio/min": ratio.min()}
  for k, v in exp.items():
    assert abs(row[k] - float(v)) <= 1e-5 + 2e-4 * abs(float(v)), (k, row[k], float(v))
  assert torch.all(info[0] == 0)
  # ---- critic loss, both variants
  for clipped in (False, True):
    values = (rets[il] + 0.3 * torch.randn(n, device=DEV)).requires_grad_(True)
    dv = torch.empty(n, device=DEV)
    ops.vf_loss(values.detach(), rets, vold, idx, dv, n, 1.0 / n, 1.0 / n, clipped, 0.2, info, slot)
    if clipped:
      vc = vold[il] + (values - vold[il]).clamp(-0.2, 0.2)
      l = 0.5 * torch.max((values - rets[il]).pow(2), (vc - rets[il]).pow(2)).mean()
    else:
      l = F.mse_loss(values, rets[il])
    l.backward()
    assert rel(dv, values.grad) < 1e-5
    assert abs(float(info[1, 4]) - float(l)) < 1e-5 * max(1, abs(float(l)))


def test_clip_adam_matches_torch_adam():
  engine, ops = _ops()
  torch.manual_seed(7)
  n = 100003
  p = torch.randn(n, device=DEV)
  ref = torch.nn.Parameter(p.clone())
  opt = torch.optim.Adam([ref], lr=3e-4, eps=1e-5)
  m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
  hyper = torch.tensor([3e-4, 0.9, 0.999, 1e-5, 0.5, 0, 0, 0], device=DEV)
  info = torch.zeros(1, 32, device=DEV)
  slot = torch.zeros(1, device=DEV, dtype=torch.int32)
  for step in range(3):
    gr = torch.randn(n, device=DEV) * (10.0 if step == 0 else 1e-4)   # clipped / unclipped
    ref.grad = gr.clone()
    norm = torch.nn.utils.clip_grad_norm_([ref], 0.5)
    opt.step()
    ops.clip_adam(p, gr, m, v, n, hyper, info, slot, 5)
    assert abs(float(info[0, 5]) - float(norm)) < 1e-4 * float(norm)
    assert rel(p, ref.data) < 1e-6
  assert float(hyper[5]) == 3.0


def test_select_rows_and_slot():
  engine, ops = _ops()
  flat = torch.arange(40, device=DEV, dtype=torch.int32) * 3
  slot = torch.zeros(1, device=DEV, dtype=torch.int32)
  cur = torch.zeros(8, device=DEV, dtype=torch.int32)
  for s in range(5):
    ops.select_rows(flat, slot, cur, 8)
    assert cur.tolist() == [3 * (8 * s + i) for i in range(8)]
    ops.slot_advance(slot, 0)
  ops.slot_advance(slot, 6)
  assert int(slot) == 0


# =================================================================================================
# round-2 entry points: minibatch prologue, epoch advantage statistics, fused optimiser tail
# =================================================================================================
def test_mb_begin_and_epoch_stats():
  engine, ops = _ops()
  torch.manual_seed(3)
  N, n, n_mb, S, Sp = 4096, 1000, 4, 93, 128
  rng = np.random.default_rng(3)
  flat = torch.tensor(rng.integers(0, N, size=n_mb * n).astype(np.int32), device=DEV)
  adv = torch.randn(N, device=DEV) * 3
  state = torch.randn(N, S, device=DEV)
  slot = torch.zeros(1, device=DEV, dtype=torch.int32)
  cur = torch.zeros(n, device=DEV, dtype=torch.int32)
  stats = torch.zeros(8, device=DEV, dtype=torch.float64)
  st16 = torch.full((n, Sp), float("nan"), device=DEV, dtype=torch.float16)
  table = torch.zeros((n_mb, 8), device=DEV, dtype=torch.float64)
  ops.adv_stats_epoch(flat, n_mb, n, adv, table)
  for s in range(n_mb):
    slot.fill_(s)
    ops.mb_begin(flat, slot, cur, n, adv, stats, state, S, st16, Sp)
    rows = flat[s * n:(s + 1) * n].long()
    assert torch.equal(cur.long(), rows)
    a = adv[rows].double()
    ref = [float(a.sum()), float((a * a).sum()), float(n), float(a.max()), float(a.min())]
    for got in (stats, table[s]):
      np.testing.assert_allclose(got[:5].cpu().numpy(), ref, rtol=1e-12)
    assert torch.equal(st16[:, :S], state[rows].half()) and float(st16[:, S:].abs().max()) == 0.0
  # the same launch without statistics / without proprio rows
  ops.mb_begin(flat, slot, cur, n, adv, None)
  assert torch.equal(cur.long(), flat[(n_mb - 1) * n:].long())


def test_opt_tail_matches_clip_adam_and_pack():
  """v4l_opt_tail phase 2 (norm from the bucket) == v4l_clip_adam, and the fp16 operand copies it writes ==
  v4l_pack_f16 of the updated bucket through the same table."""
  engine, ops = _ops()
  torch.manual_seed(9)
  n = 50000                                   # multiple of 4 (buckets are padded to 16 bytes)
  p0 = torch.randn(n, device=DEV)
  rng = np.random.default_rng(9)
  # packing table: every parameter appears at two positions of a padded fp16 buffer, some entries are padding
  n_pack = 2 * n + 1000
  perm = rng.permutation(n_pack)
  table = -np.ones(n_pack, np.int32)
  table[perm[:n]] = np.arange(n); table[perm[n:2 * n]] = np.arange(n)
  scatter = -np.ones((n, 4), np.int32)
  scatter[:, 0] = perm[:n]; scatter[:, 1] = perm[n:2 * n]
  scatter[::7, 2] = perm[:n][::7]             # some parameters also live in the "other" network's buffer
  table_t, scatter_t = torch.tensor(table, device=DEV), torch.tensor(scatter, device=DEV)
  for step in range(3):
    gr = torch.randn(n, device=DEV) * (10.0 if step == 0 else 1e-4)
    if step == 0:
      pa, pb = p0.clone(), p0.clone()
      ma, va, mb_, vb = (torch.zeros(n, device=DEV) for _ in range(4))
      ha = torch.tensor([3e-4, 0.9, 0.999, 1e-5, 0.5, 0, 0, 0], device=DEV); hb = ha.clone()
      ia, ib = torch.zeros(1, 32, device=DEV), torch.zeros(1, 32, device=DEV)
      slot = torch.zeros(1, device=DEV, dtype=torch.int32)
      adv_slot = torch.zeros(1, device=DEV, dtype=torch.int32)
      packed_self = torch.zeros(n_pack, device=DEV, dtype=torch.float16)
      packed_other = torch.zeros(n_pack, device=DEV, dtype=torch.float16)
    ops.clip_adam(pa, gr, ma, va, n, ha, ia, slot, 5)
    ops.opt_tail(2, param=pb, grad=gr, m=mb_, v=vb, n=n, hyper=hb, info=ib, slot=slot, norm_slot=5,
                 scatter=scatter_t, packed_self=packed_self, packed_other=packed_other, slot_advance=adv_slot)
    torch.cuda.synchronize()
    assert abs(float(ia[0, 5]) - float(ib[0, 5])) <= 1e-6 * float(ia[0, 5])
    assert rel(pb, pa) < 1e-6 and rel(mb_, ma) < 1e-6 and rel(vb, va) < 1e-6
    ref = torch.zeros(n_pack, device=DEV, dtype=torch.float16)
    ops.pack_f16(pb, table_t, ref, n_pack)
    assert torch.equal(packed_self, ref)
    assert torch.equal(packed_other[perm[:n][::7]], pb[::7].half())
  assert float(hb[5]) == 3.0 and int(adv_slot) == 3


def test_pf_loss_stats_table_and_f16_gradient():
  """the per-minibatch statistics table (stats_per_slot) and the loss-scaled fp16 gradient row"""
  engine, ops = _ops()
  torch.manual_seed(4)
  n, A = 700, 12
  mean, tmean, acts = torch.randn(n, A, device=DEV) * 0.1, torch.randn(n, A, device=DEV) * 0.1, torch.randn(n, A, device=DEV) * 0.2
  logstd = torch.full((A,), math.log(0.125), device=DEV); adv = torch.randn(n, device=DEV)
  st = torch.zeros((3, 8), device=DEV, dtype=torch.float64)
  a = adv.double()
  st[2, 0], st[2, 1], st[2, 2], st[2, 3], st[2, 4] = a.sum(), (a * a).sum(), n, a.max(), a.min()
  slot = torch.full((1,), 2, device=DEV, dtype=torch.int32)
  info = torch.zeros(3, 32, device=DEV)
  outs = []
  for per_slot in (False, True):
    d_mean = torch.zeros(n, A, device=DEV); d_ls = torch.zeros(A, device=DEV)
    d16 = torch.full((n, 16), float("nan"), device=DEV, dtype=torch.float16)
    ops.pf_loss(mean, logstd, tmean, logstd, acts, adv, None, st if per_slot else st[2], d_mean, d_ls, n, A, 1.0 / n, 1.0 / n,
                0.2, 0.005, info, slot, d_f16=d16, scale_f16=256.0, stats_per_slot=per_slot)
    torch.cuda.synchronize()
    outs.append((d_mean.clone(), d_ls.clone(), info[2].clone()))
    assert torch.equal(d16[:, :A], (d_mean * 256.0).half()) and float(d16[:, A:].abs().max()) == 0.0
  for x, y in zip(outs[0], outs[1]):
    assert torch.equal(x, y)
