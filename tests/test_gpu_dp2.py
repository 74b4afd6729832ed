"""Data-parallel parity on TWO GPUs (NCCL): a world-2 run in which each rank owns half of the env columns
and half of every minibatch must reproduce the world-1 run on the whole rollout — same time-row permutations,
global advantage statistics, gradient buckets SUM-all-reduced once per optimiser step, loss means over the
GLOBAL minibatch (reference minibatch structure on_policy.py:76-89).  This exercises the product's
own DP path: NCCL collectives captured inside the per-minibatch CUDA graph, `inv_global` scaling of both losses
and of the entropy bonus' logstd gradient, the split reduce -> all-reduce -> clip+Adam optimiser tail.

Exact tier: parameters within 1e-5 (only the summation order of the batch reductions differs).
Tensor-core tier: within the fp16 tier's single-step noise (the two runs round the same tensors identically, so
they agree far better than either agrees with fp32).
Skipped when fewer than two CUDA devices are visible.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _run(rank, world, port, precision, family, q):
  try:
    import torch.distributed as dist
    from benchutil import synth
    from benchutil.harness import build_nets, load_np_sd, fill_buffer, make_ppo
    from tests import _golden as g
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    pg = None
    if world > 1:
      os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
      import datetime
      dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
      pg = dist.group.WORLD
    S, A = g.FAMILIES[family]
    T, E, B = 8, 4, 16
    roll = synth.make_rollout(4321, T, E, S, A, with_img=family != "mlp", p_term=0.15, time_limit_p=0.1)
    El = E // world
    sl = slice(rank * El, (rank + 1) * El)
    mine = {k: (v[:, sl] if (v.ndim == 3 and v.shape[1] == E) else v) for k, v in roll.items()}
    mine["last_obs"] = roll["last_obs"][sl]
    mine["last_terminals"] = roll["last_terminals"][sl]
    pf, vf = build_nets(family, S, A)
    pf_np, vf_np = g.family_weights(family)
    load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
    pf, vf = pf.to(dev), vf.to(dev)
    buf = fill_buffer(mine, T, El)
    agent, logger = make_ppo(pf, vf, buf, A, B // world, T * El, 2, device=dev)
    agent.process_group = pg
    agent.precision = precision
    for epoch in (3, 4):
      agent.current_epoch = epoch
      np.random.seed(100 + epoch)              # same time-row permutations on every rank
      agent.update_per_epoch()
    torch.cuda.synchronize(dev)
    out = {"pf": {k: v.detach().cpu().numpy() for k, v in pf.state_dict().items()},
           "vf": {k: v.detach().cpu().numpy() for k, v in vf.state_dict().items()},
           "vf_loss": [i["Training/vf_loss"] for i in logger.infos],
           "policy_loss": [i["Training/policy_loss"] for i in logger.infos]}
    q.put((world, rank, out))
    q.close(); q.join_thread()               # flush the feeder thread: os._exit below would drop the message
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize(dev)
    os._exit(0)      # CUDA graphs still reference the communicator: skip NCCL teardown
  except Exception:
    import traceback
    q.put((world, rank, traceback.format_exc()))
    q.close(); q.join_thread()
    os._exit(1)


def _launch(world, precision, family):
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_run, args=(r, world, port, precision, family, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = []
  try:
    for _ in procs:
      item = q.get(timeout=240)              # a rank that died takes its peer down with it: never wait long
      res.append(item)
      if not isinstance(item[2], dict):
        break
  finally:
    for p in procs:
      p.join(timeout=30 if len(res) == len(procs) and all(isinstance(i[2], dict) for i in res) else 1)
      if p.is_alive():
        p.kill()
  for w, r, out in res:
    assert isinstance(out, dict), "world %d rank %d failed:\n%s" % (w, r, out)
  assert len(res) == len(procs)
  return {r: out for _, r, out in res}


def _rel(a, b):
  return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


@pytest.mark.parametrize("precision,family,tol", [("f16", "loco", 5e-3), ("fp32", "loco", 1e-5), ("fp32", "mlp", 1e-5)])
def test_dp_world2_matches_world1(precision, family, tol):
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two CUDA devices")
  one = _launch(1, precision, family)[0]
  two = _launch(2, precision, family)
  worst = 0.0
  for net in ("pf", "vf"):
    for k, ref in one[net].items():
      for r in (0, 1):                       # every rank holds the same parameters
        worst = max(worst, _rel(two[r][net][k], ref))
  print("DP world-2 vs world-1 (%s, %s): worst relative parameter difference %.3e" % (precision, family, worst))
  assert worst < tol
  # both ranks step identically (bit-identical parameters): the all-reduced buckets and the clip factor agree
  for net in ("pf", "vf"):
    for k in one[net]:
      assert np.array_equal(two[0][net][k], two[1][net][k]), (net, k)
  # logged losses are LOCAL means (reduced once per epoch by the caller): their rank average is the global mean
  for key in ("vf_loss", "policy_loss"):
    avg = (np.asarray(two[0][key]) + np.asarray(two[1][key])) / 2
    assert np.allclose(avg, one[key], rtol=max(tol * 20, 1e-4), atol=1e-5), (key, avg, one[key])
