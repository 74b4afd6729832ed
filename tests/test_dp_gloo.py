"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic of the PPO engine:
  * the global advantage statistics assembled by PPOUpdateEngine._allreduce_stats,
  * the gradient identity the engine relies on: sum over ranks of local-minibatch gradients taken
    with loss scale 1/B_global and GLOBAL advantage normalisation == full-minibatch gradient
    (checked with the CPU oracle as the per-rank compute).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import _golden as g


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from oracle import ppo_oracle as po, synth
    from vision4leg_b200.algo.on_policy.ppo_engine import PPOUpdateEngine
    torch.set_num_threads(2)
    family, (S, A) = "mlp", g.FAMILIES["mlp"]
    B = 24
    rng = np.random.default_rng(0)
    roll = synth.make_rollout(3, B, 1, S, A, with_img=False)
    obs, acts = roll["obs"].reshape(B, -1), roll["acts"].reshape(B, -1)
    advs = rng.standard_normal((B, 1)).astype(np.float32)
    rets = rng.standard_normal((B, 1)).astype(np.float32)
    lo, hi = rank * B // world, (rank + 1) * B // world

    # ---- (1) global advantage statistics through the engine's own combine code
    class Fake:
      pass
    eng = Fake()
    eng.pg, eng.world = dist.group.WORLD, world
    a = advs[lo:hi, 0].astype(np.float64)
    b = {"stats": torch.tensor([a.sum(), (a * a).sum(), len(a), a.max(), a.min(), 0, 0, 0], dtype=torch.float64),
         "stats_all": torch.zeros((world, 8), dtype=torch.float64)}
    PPOUpdateEngine._allreduce_stats(eng, b)
    full = advs[:, 0].astype(np.float64)
    want = [full.sum(), (full * full).sum(), B, full.max(), full.min()]
    np.testing.assert_allclose(b["stats"][:5].numpy(), want, rtol=1e-12)
    n = b["stats"][2].item()
    mean = b["stats"][0].item() / n
    std = np.sqrt((b["stats"][1].item() - b["stats"][0].item() ** 2 / n) / (n - 1))
    np.testing.assert_allclose([mean, std], [full.mean(), full.std(ddof=1)], rtol=1e-10)

    # ---- (2) gradient identity
    pf_np, vf_np = g.family_weights(family)
    pf, vf = po.sd_to_torch(pf_np, vf_np)
    names_v = list(vf.keys())
    names_p = list(pf.keys())

    def grads(sl, scale, adv_mean, adv_std):
      o = torch.tensor(obs[sl]); ac = torch.tensor(acts[sl])
      pv = [vf[k].clone().requires_grad_(True) for k in names_v]
      Pv = dict(zip(names_v, pv))
      lv = ((po.mlp_forward(Pv, o) - torch.tensor(rets[sl])) ** 2).sum() * scale
      gv = torch.autograd.grad(lv, pv)
      pp = [pf[k].clone().requires_grad_(True) for k in names_p]
      Pp = dict(zip(names_p, pp))
      lp, ent, _ = po.gaussian_update(po.mlp_forward(Pp, o), Pp["logstd"], ac)
      with torch.no_grad():
        tlp, _, _ = po.gaussian_update(po.mlp_forward(pf, o) + 0.01, pf["logstd"], ac)
      ah = (torch.tensor(advs[sl]) - adv_mean) / (adv_std + 1e-5)
      ratio = torch.exp(lp - tlp)
      loss = (-torch.min(torch.clamp(ratio, 0.8, 1.2) * ah, ratio * ah)).sum() * scale \
        - 0.005 * ent.sum() * scale
      gp = torch.autograd.grad(loss, pp)
      return torch.cat([x.reshape(-1) for x in gv]), torch.cat([x.reshape(-1) for x in gp])

    gv, gp = grads(slice(lo, hi), 1.0 / B, float(mean), float(std))
    dist.all_reduce(gv); dist.all_reduce(gp)
    fv, fp = grads(slice(0, B), 1.0 / B, float(full.mean()), float(full.std(ddof=1)))
    assert float((gv - fv).abs().max() / fv.abs().max()) < 1e-5
    assert float((gp - fp).abs().max() / fp.abs().max()) < 1e-5
    q.put((rank, "ok"))
  except Exception as e:   # surface the failure to the parent
    import traceback
    q.put((rank, traceback.format_exc()))
  finally:
    dist.destroy_process_group()


def test_data_parallel_host_logic_world2():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = [q.get(timeout=240) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  for rank, msg in results:
    assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_entropy_bonus_share_formula():
  """The actor-loss kernel writes d_logstd = s_local - c * n_local * inv_global on every rank (csrc/ppo_ops.cu,
  pf_loss finalize) and the buckets are SUM-all-reduced: the entropy bonus' gradient must come out as -c once,
  not once per rank (round-1 advisor finding: it was -c * world)."""
  c, world, n_local = 0.005, 8, 128
  inv_global = 1.0 / (n_local * world)
  s_local = np.random.default_rng(0).standard_normal(world)
  summed = sum(s - c * n_local * inv_global for s in s_local)
  assert abs(summed - (s_local.sum() - c)) < 1e-12


def _norm_worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from oracle import obs_oracle as oo
    from vision4leg_b200 import obs_pipeline as op
    S = 23
    rng = np.random.RandomState(17)
    ref = oo.Normalizer((S,))
    mean, var, count = torch.zeros(S, dtype=torch.float64), torch.ones(S, dtype=torch.float64), 1e-4
    for step, sizes in enumerate([(3, 5), (4, 4), (0, 6), (7, 1)]):            # rows per rank: uneven, one rank empty once
      rows = [(rng.randn(n, S) * (1 + 10 * step) + step).astype(np.float32) for n in sizes]
      mine = rows[rank].astype(np.float64)
      # what kernel mode 2 produces on a rank: batch mean / population variance of ITS rows
      bm = torch.from_numpy(mine.mean(0) if len(mine) else np.zeros(S))
      bv = torch.from_numpy(mine.var(0) if len(mine) else np.zeros(S))
      gm, gv, gn = op.gather_batch_stats(bm, bv, len(mine), dist.group.WORLD)
      mean, var, count = op.merge_mean_var_count(mean, var, count, gm, gv, gn)
      ref.update(np.concatenate(rows).astype(np.float64))                        # one process over the union of the rows
      np.testing.assert_allclose(mean.numpy(), ref.mean, rtol=1e-12, atol=1e-13)
      np.testing.assert_allclose(var.numpy(), ref.var, rtol=1e-12, atol=1e-13)
      assert abs(count - ref.count) < 1e-9
    q.put((rank, "ok"))
  except Exception:
    import traceback
    q.put((rank, traceback.format_exc()))
  finally:
    dist.destroy_process_group()


def test_data_parallel_normalizer_statistics_world2():
  """the observation normaliser under data parallelism: per-rank batch statistics all-gathered and combined equal
  the statistics of ONE process over the union of the ranks' rows (reference base_wrapper.py:44-61,76-86)"""
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_norm_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = [q.get(timeout=240) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  for rank, msg in results:
    assert msg == "ok", "rank %d: %s" % (rank, msg)
