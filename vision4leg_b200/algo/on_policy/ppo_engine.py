"""Device-resident PPO update engine: everything `PPO.update_per_epoch` does after the rollout
reaches the GPU (reference torchrl/algo/on_policy/ppo.py:28-153).

Layout in HBM
  * parameters: ONE flat fp32 bucket  [ pf-only | shared encoder | vf-only ]  (16-byte aligned
    segments); the modules' nn.Parameters are views into it, so checkpoints are unchanged.  The
    actor's optimiser bucket is the prefix, the critic's the suffix — each is one contiguous
    range for the fused clip+Adam kernel and for the (single) gradient all-reduce, and the
    shared encoder is stepped by BOTH with separate Adam moments.
  * rollout: aligned planes state[N,S], img[N,4*64*64], acts[N,A], values/advs/returns[N]
    (the host buffer's rows are split by a strided H2D copy; SURVEY §7 hard part 7).
  * a minibatch is a row-index list (time rows x envs, reference on_policy.py:73-92) consumed
    by the first-layer gathers — no minibatch copy is ever materialised.

One minibatch = the fixed kernel sequence of `_minibatch`; it is captured once in a CUDA graph
and replayed (the minibatch number lives in a device-side slot counter).  Logged statistics
accumulate in a device info table read back once per epoch.
"""
import numpy as np
import torch

from ... import engine
from ... import engine_tc
from ..._lib import INFO_KEYS, INFO_STRIDE, INFO_GRAD_NORM_PF, INFO_GRAD_NORM_VF, V4LError

_ALIGN = 4   # floats (16 B)
GRAPH_GROUP = 4   # minibatches per CUDA-graph replay in the later opt-epochs


def _round_up(n, a):
  return (n + a - 1) // a * a


class _Bucket:
  """Flat parameter storage with named views."""

  def __init__(self, named_groups, device):
    # named_groups: [(group_name, [(name, Parameter)])] in layout order
    self.offsets = {}
    self.group_range = {}
    off = 0
    for gname, items in named_groups:
      start = off
      for name, p in items:
        self.offsets[id(p)] = (off, p.numel(), tuple(p.shape))
        off += _round_up(p.numel(), _ALIGN)
      self.group_range[gname] = (start, off)
    self.size = off
    self.flat = torch.zeros(off, device=device, dtype=torch.float32)
    for gname, items in named_groups:
      for name, p in items:
        o, n, shape = self.offsets[id(p)]
        view = self.flat[o:o + n].view(shape)
        view.copy_(p.data)
        p.data = view

  def view_of(self, flat, p, base=0):
    o, n, shape = self.offsets[id(p)]
    return flat[o - base:o - base + n].view(shape)


class PPOUpdateEngine:
  def __init__(self, pf, vf, target_pf, device, clip_para, entropy_coeff, clipped_value_loss,
               use_cuda_graph=True, process_group=None, precision="fp32"):
    self.device = torch.device(device)
    if self.device.type != "cuda":
      raise V4LError("the PPO update runs on a CUDA device (got %s); there is no CPU fallback"
                     % (self.device,))
    if getattr(pf, "tanh_action", False):
      raise NotImplementedError("tanh_action policies are not implemented on the CUDA PPO path "
                                "(no shipped config uses them)")
    self.ops = engine.ops_for(self.device)
    self.device = self.ops.device
    self.pf, self.vf, self.target_pf = pf, vf, target_pf
    self.family = pf._family
    if vf._family != self.family:
      raise V4LError("pf and vf must be the same network family")
    self.S = pf._state_dim
    self.A = pf._out_dim
    self.has_img = pf._has_img
    self.clip_para = float(clip_para)
    self.entropy_coeff = float(entropy_coeff)
    self.clipped_value_loss = bool(clipped_value_loss)
    self.use_cuda_graph = use_cuda_graph
    self.pg = process_group
    self.world = 1
    if process_group is not None:
      import torch.distributed as dist
      self.world = dist.get_world_size(process_group)
    self._build_buckets()
    kw = getattr(pf, "_plan_kwargs", {})
    if precision == "fp16":
      precision = "f16"
    if precision not in ("fp32", "f16"):
      raise ValueError("precision must be 'fp32' (exact CUDA-core tier) or 'f16' (tcgen05 tier)")
    self.precision = precision
    if precision == "f16":
      if self.family not in engine_tc.PLANS:
        raise NotImplementedError("the tensor-core tier covers the image families (LocoTransformer, NatureCNN and "
                                  "their vision-only variants); use precision='fp32' for %s" % self.family)
      nh = kw.get("n_heads", (1, 1))
      Plan = engine_tc.PLANS[self.family]
      self.plan_pf = Plan(self.ops, self.S, self.A, self.pf_layout, nh)
      self.plan_vf = Plan(self.ops, self.S, 1, self.vf_layout, nh)
      self.plan_t = Plan(self.ops, self.S, self.A, self.pf_layout, nh, with_backward=False)
      self.plan_t.pack(self.t_flat)
      self.plan_pf.world = self.plan_vf.world = self.world
      self._build_scatter()
    else:
      self.plan_pf = engine.make_plan(self.family, self.ops, self.S, self.A, **kw)
      self.plan_vf = engine.make_plan(self.family, self.ops, self.S, 1, **kw)
      self.plan_t = engine.make_plan(self.family, self.ops, self.S, self.A, **kw)
    self._graphs = {}
    self._roll = None
    self._mb_bufs = {}
    self._epoch_stats = None          # [minibatches, 8] doubles while an epoch of uniform minibatches runs

  # ---------------------------------------------------------------------------------------------
  def _build_buckets(self):
    dev = self.device
    pf_named = list(self.pf.named_parameters())
    vf_named = list(self.vf.named_parameters())
    vf_ids = {id(p) for _, p in vf_named}
    pf_ids = {id(p) for _, p in pf_named}
    pf_only = [(n, p) for n, p in pf_named if id(p) not in vf_ids]
    shared = [(n, p) for n, p in pf_named if id(p) in vf_ids]
    vf_only = [(n, p) for n, p in vf_named if id(p) not in pf_ids]
    for _, p in pf_named + vf_named:
      if p.device != dev or p.dtype != torch.float32:
        raise V4LError("all parameters must be fp32 on %s before the PPO engine is built" % dev)
    self.bucket = b = _Bucket([("pf_only", pf_only), ("shared", shared), ("vf_only", vf_only)], dev)
    self.pf_range = (b.group_range["pf_only"][0], b.group_range["shared"][1])
    self.vf_range = (b.group_range["shared"][0], b.group_range["vf_only"][1])
    self.n_pf = self.pf_range[1] - self.pf_range[0]
    self.n_vf = self.vf_range[1] - self.vf_range[0]
    self.pf_flat = b.flat[self.pf_range[0]:self.pf_range[1]]
    self.vf_flat = b.flat[self.vf_range[0]:self.vf_range[1]]
    z = lambda n: torch.zeros(n, device=dev, dtype=torch.float32)
    self.g_pf, self.g_vf = z(self.n_pf), z(self.n_vf)
    self.m_pf, self.v_pf, self.m_vf, self.v_vf = z(self.n_pf), z(self.n_pf), z(self.n_vf), z(self.n_vf)
    hyper = lambda: torch.tensor([0.0, 0.9, 0.999, 1e-5, 0.5, 0.0, 0.0, 0.0], device=dev)
    self.hyper_pf, self.hyper_vf = hyper(), hyper()
    # name -> tensor dicts the plans consume (logstd is not a network weight)
    self.P_pf = {n: p.data for n, p in pf_named if n != "logstd"}
    self.P_vf = {n: p.data for n, p in vf_named}
    self.G_pf = {n: b.view_of(self.g_pf, p, self.pf_range[0]) for n, p in pf_named}
    self.G_vf = {n: b.view_of(self.g_vf, p, self.vf_range[0]) for n, p in vf_named}
    self.logstd = dict(pf_named)["logstd"].data
    # bucket-relative layouts {name: (offset, shape)} for the tensor-core tier's packing tables
    self.pf_layout = {n: (b.offsets[id(p)][0] - self.pf_range[0], tuple(p.shape)) for n, p in pf_named}
    self.vf_layout = {n: (b.offsets[id(p)][0] - self.vf_range[0], tuple(p.shape)) for n, p in vf_named}
    # frozen target policy: own flat copy in the actor-bucket layout
    t_named = list(self.target_pf.named_parameters())
    assert [n for n, _ in t_named] == [n for n, _ in pf_named]
    self.t_flat = torch.zeros(self.n_pf, device=dev, dtype=torch.float32)
    self.P_t = {}
    for (n, tp), (_, p) in zip(t_named, pf_named):
      view = b.view_of(self.t_flat, p, self.pf_range[0])
      view.copy_(tp.data)
      tp.data = view
      if n != "logstd":
        self.P_t[n] = view
    self.t_logstd = dict(t_named)["logstd"].data
    self._param_ptrs = [(p, p.data_ptr()) for _, p in pf_named + vf_named + t_named]

  def _build_scatter(self):
    """For every element of an optimiser bucket: where its fp16 copies live in the packed operand buffers
    (TcWeights: tap-major forward + data-gradient orientation) of its own network and — shared-encoder
    weights — of the other one.  The optimiser tail writes them in the Adam pass (v4l_opt_tail)."""
    def positions(W, lo_src, lo_dst, n_dst):
      tab = W.table.cpu().numpy().astype(np.int64)
      pos = np.nonzero(tab >= 0)[0]
      dst = tab[pos] + lo_src - lo_dst              # index in the destination bucket
      ok = (dst >= 0) & (dst < n_dst)
      pos, dst = pos[ok], dst[ok]
      out = -np.ones((n_dst, 2), np.int64)
      order = np.argsort(dst, kind="stable")
      pos, dst = pos[order], dst[order]
      first = np.r_[True, dst[1:] != dst[:-1]]
      rank = np.arange(len(dst)) - np.maximum.accumulate(np.where(first, np.arange(len(dst)), 0))
      assert rank.max(initial=0) <= 1, "a parameter appears in more than two packed positions"
      out[dst, rank] = pos
      return out
    self.scatter = {}
    for name, (lo, n, W_self, lo_other, W_other) in {
        "vf": (self.vf_range[0], self.n_vf, self.plan_vf.W, self.pf_range[0], self.plan_pf.W),
        "pf": (self.pf_range[0], self.n_pf, self.plan_pf.W, self.vf_range[0], self.plan_vf.W)}.items():
      t = np.concatenate([positions(W_self, lo, lo, n), positions(W_other, lo_other, lo, n)], axis=1)
      self.scatter[name] = torch.tensor(t.astype(np.int32), device=self.device).contiguous()

  def check_views(self):
    for p, ptr in self._param_ptrs:
      if p.data_ptr() != ptr:
        raise V4LError("a network parameter was re-allocated (e.g. .to()/.half()) after the PPO "
                       "engine flattened it; rebuild the algorithm object")

  # ---------------------------------------------------------------------------------------------
  # rollout store
  # ---------------------------------------------------------------------------------------------
  def _alloc_rollout(self, T, E):
    N, dev = T * E, self.device
    f = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    r = dict(T=T, E=E, N=N, state=f(N, self.S), acts=f(N, self.A), values=f(N), rewards=f(N),
             terminals=f(N), advs=f(N), rets=f(N), last_value=f(E))
    if self.has_img:
      if self.precision == "f16":
        # the tensor-core tier keeps only the fp16 space-to-depth image (32 KB / transition: a 2^20-transition
        # sweep is 34 GB); fp32 rows pass through a bounded staging chunk on the way in (_copy_obs_rows)
        r["imgs"] = torch.empty((N, 16, 16, 64), device=dev, dtype=torch.float16)
      else:
        r["img"] = f(N, engine.IMG_ELEMS)
    if self.precision == "f16":
      r["tmean_all"] = f(N, self.A)        # target-policy action means, filled during the first opt-epoch
    self._roll = r
    self._graphs.clear()            # captured graphs hold the old planes' addresses
    return r

  def load_rollout(self, buf, stream_obs=False):
    """Pinned host rows -> device planes (async on the current stream).  With stream_obs the
    observation rows are NOT copied here: run_epoch streams them on a copy stream in the order the
    first opt-epoch consumes them, so the 1 GB host->device transfer overlaps the first
    minibatches instead of preceding them."""
    host = buf._host
    T, E = buf._max_replay_buffer_size, buf.env_nums
    r = self._roll
    if r is None or r["T"] != T or r["E"] != E:
      r = self._alloc_rollout(T, E)
    obs = host["obs"]
    D = obs.shape[-1]
    expect = self.S + (engine.IMG_ELEMS if self.has_img else 0)
    if D != expect:
      raise V4LError("rollout observation width %d, expected %d" % (D, expect))
    self._pending_obs = None
    self._pending_half = None
    resident = getattr(self, "_resident_rows", None)
    self._resident_rows = set()
    if resident is not None and len(resident) == T and self.precision == "f16":
      # every time row of this epoch was written on the device by act(row=t): no observation H2D at all
      self.h2d_bytes = 0
      stream_obs, skip_obs = False, True
    else:
      skip_obs = False
    if skip_obs:
      pass
    elif stream_obs:
      self._pending_obs = (obs, D)
      if self.precision == "f16" and self.has_img and host.get("obs_img16") is not None \
          and getattr(buf, "_half_S", None) == self.S:
        # stream the buffer's fp16 staging copy of the depth stack (+ packed proprio rows) instead
        self._pending_half = (host["obs_img16"], host["obs_state"])
    else:
      self._copy_obs_rows(obs, D, 0, T * E)
    self.h2d_bytes = 0 if skip_obs else T * E * D * 4
    if self._pending_half is not None:
      self.h2d_bytes = T * E * (engine.IMG_ELEMS * 2 + self.S * 4)
    for key in ("acts", "values", "rewards", "terminals"):
      src = host[key].reshape(T * E, -1)
      r[key].view(T * E, -1).copy_(src, non_blocking=True)
      self.h2d_bytes += src.numel() * 4
    tl = host.get("time_limits")
    if tl is not None:
      r["time_limits"] = tl.reshape(T, -1).to(self.device, non_blocking=True)
      self.h2d_bytes += tl.numel() * 4
    else:
      r["time_limits"] = None
    return r

  def _copy_obs_rows(self, obs, D, n0, n):
    """rows [n0, n0+n) of the pinned [N, D] host observations -> state / image planes"""
    r = self._roll
    base = obs.data_ptr() + n0 * D * 4
    if self.S:
      self.ops.h2d_2d(r["state"][n0:], self.S * 4, base, D * 4, self.S * 4, n)
    if self.has_img and self.precision == "f16":
      # fp32 CHW rows -> bounded staging chunk -> fp16 space-to-depth NHWC (v4l_ingest_img)
      chunk = min(n, 4096)
      stage = r.get("img_stage")
      if stage is None or stage.shape[0] < chunk:
        stage = r["img_stage"] = torch.empty((chunk, engine.IMG_ELEMS), device=self.device, dtype=torch.float32)
      for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        self.ops.h2d_2d(stage, engine.IMG_ELEMS * 4, base + (c0 * D + self.S) * 4, D * 4, engine.IMG_ELEMS * 4, m)
        self.ops.ingest_img(stage, r["imgs"][n0 + c0:], m)
    elif self.has_img:
      self.ops.h2d_2d(r["img"][n0:], engine.IMG_ELEMS * 4, base + self.S * 4, D * 4, engine.IMG_ELEMS * 4, n)

  def _stream_chunk(self, obs, D, trows, E, k, rows):
    """Copy the observation rows of minibatch k (first opt-epoch) on the copy stream, convert
    them for the tensor-core tier, and return the event that marks them resident."""
    r = self._roll
    # free-running copy stream: it was ordered after the main stream ONCE (run_epoch); waiting for
    # the main stream here would serialise copy k+1 behind minibatch k
    with self.ops.fork(2, wait=False):
      half = getattr(self, "_pending_half_cur", None)
      if half is not None:
        img16, st32 = half
        stage = r.get("stage16")
        if stage is None:
          stage = r["stage16"] = torch.empty((r["N"], engine.IMG_ELEMS), device=self.device, dtype=torch.float16)
        tr = np.ascontiguousarray(trows, dtype=np.int32)
        self.ops.h2d_rows(stage.data_ptr(), img16.data_ptr(), tr, E * engine.IMG_ELEMS * 2)
        if self.S:        # proprio rows go straight into their fp32 plane (same layout on both sides)
          assert st32.shape[-1] == self.S
          self.ops.h2d_rows(r["state"].data_ptr(), st32.data_ptr(), tr, E * self.S * 4)
        self.ops.ingest_img_f16(stage.data_ptr(), r["imgs"], len(trows) * E, self._flat_idx[k * rows * E:])
      elif self.has_img:
        # copy engine: one contiguous 8-row block (E x D floats) per time row into a device staging
        # matrix with the host layout; then ONE kernel splits/convert this minibatch's rows into the
        # device layouts (fp32 image plane only for the exact tier).  No SM is tied up waiting on PCIe.
        stage = r.get("stage")
        if stage is None:
          stage = r["stage"] = torch.empty((r["N"], D), device=self.device, dtype=torch.float32)
        self.ops.h2d_rows(stage.data_ptr(), obs.data_ptr(), np.ascontiguousarray(trows, dtype=np.int32), E * D * 4)
        self.ops.ingest_rows(stage.data_ptr(), D, self.S, self._flat_idx[k * rows * E:], len(trows) * E,
                             r["state"] if self.S else None,
                             r["img"] if self.precision != "f16" else None,
                             r.get("imgs") if self.precision == "f16" else None)
      else:
        for t in trows:
          self._copy_obs_rows(obs, D, int(t) * E, E)
      ev = torch.cuda.Event()
      ev.record()
    return ev

  def load_rollout_arrays(self, roll):
    """Test/bench helper: same as load_rollout from a dict of [T,E,*] numpy arrays."""
    class _B:
      pass
    b = _B()
    T, E = roll["rewards"].shape[:2]
    b._max_replay_buffer_size, b.env_nums = T, E
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory()
    b._host = {k: pin(roll[k]) for k in ("obs", "acts", "values", "rewards", "terminals", "time_limits")
               if k in roll}
    self._pinned_keepalive = b._host
    return self.load_rollout(b)

  def compute_advantages(self, last_obs, last_terminals, gamma, tau, time_limit_filter, use_gae=True):
    """process_epoch_samples on the device (reference on_rl_algo.py:23-34)."""
    r = self._roll
    T, E = r["T"], r["E"]
    x = torch.as_tensor(np.ascontiguousarray(last_obs, dtype=np.float32)).to(self.device).reshape(E, -1)
    v = torch.empty((E, 1), device=self.device, dtype=torch.float32)
    plan = self._aux_plan(E)
    if self.precision == "f16":
      imgs = torch.empty((E, 16, 16, 64), device=self.device, dtype=torch.float16)
      self.ops.ingest_img(x[:, self.S:].clone(memory_format=torch.contiguous_format), imgs, E)
      st = torch.empty((E, plan.Sp), device=self.device, dtype=torch.float16)
      self.ops.gather_rows_f16(x, True, None, st, E, self.S, x.shape[1], plan.Sp)
      plan.pack(self.vf_flat)
      plan.forward(self.vf_flat, imgs, None, st, E, v)
    else:
      plan.forward(self.P_vf, engine.Input.from_flat(x, self.S, self.has_img), v)
    notdone = 1.0 - torch.as_tensor(np.asarray(last_terminals, np.float32).reshape(E)).to(self.device)
    r["last_value"].copy_(v.view(E) * notdone)
    tl = r["time_limits"]
    use_tl = bool(time_limit_filter and tl is not None)
    tl_st, tl_se = (tl.shape[1], 1) if (tl is not None and tl.shape[1] == E and E > 1) else (1, 0)
    self.ops.gae(r["rewards"], r["values"], r["terminals"], tl, tl_st, tl_se, r["last_value"],
                 r["advs"], r["rets"], T, E, float(gamma), float(tau if use_gae else 1.0), use_tl,
                 0 if use_gae else 1)

  def _aux_plan(self, B, which="vf"):
    """forward-only plan for `B` rows of the critic ("vf") or the actor ("pf")"""
    key = ("aux", B, which)
    p = self._mb_bufs.get(key)
    if p is None:
      kw = getattr(self.pf, "_plan_kwargs", {})
      out_dim, layout = (1, self.vf_layout) if which == "vf" else (self.A, self.pf_layout)
      if self.precision == "f16":
        p = engine_tc.PLANS[self.family](self.ops, self.S, out_dim, layout, kw.get("n_heads", (1, 1)),
                                         with_backward=False)
      else:
        p = engine.make_plan(self.family, self.ops, self.S, out_dim, **kw)
      self._mb_bufs[key] = p
    return p

  def infer(self, obs, imgs_out=None, state_out=None):
    """Action means and values of `obs` [n, D] (fp32 host array or device tensor) on the engine's
    precision tier with the CURRENT parameters: what the collector's `pf.explore` + `vf` evaluate
    (reference collector/on_policy.py:90-100), one call for both networks.  On the tensor-core tier the
    shared encoder (conv trunk + proprio branch: 7.6 of 11 MFLOP) runs ONCE and feeds both networks — legal
    here because no optimiser step intervenes between the two forwards.  imgs_out / state_out: optional device
    destinations (rollout plane rows) for the converted observation.  Returns device tensors
    (mean [n, A], value [n, 1])."""
    x = torch.as_tensor(np.ascontiguousarray(obs, dtype=np.float32) if not torch.is_tensor(obs) else obs)
    x = x.to(self.device, torch.float32).reshape(-1, self.S + (engine.IMG_ELEMS if self.has_img else 0)).contiguous()
    n = x.shape[0]
    mean = torch.empty((n, self.A), device=self.device, dtype=torch.float32)
    value = torch.empty((n, 1), device=self.device, dtype=torch.float32)
    ppf, pvf = self._aux_plan(n, "pf"), self._aux_plan(n, "vf")
    if self.precision == "f16":
      imgs = imgs_out if imgs_out is not None else torch.empty((n, 16, 16, 64), device=self.device, dtype=torch.float16)
      self.ops.ingest_img(x[:, self.S:].clone(memory_format=torch.contiguous_format), imgs, n)
      st = torch.zeros((n, ppf.Sp), device=self.device, dtype=torch.float16)
      if self.S:
        self.ops.gather_rows_f16(x, True, None, st, n, self.S, x.shape[1], ppf.Sp)
        if state_out is not None:
          state_out.copy_(x[:, :self.S])
      ppf.pack(self.pf_flat)
      ppf.forward(self.pf_flat, imgs, None, st, n, mean)
      pvf.pack(self.vf_flat)
      pvf.forward(self.vf_flat, imgs, None, st, n, value, enc_from=ppf)
    else:
      inp = engine.Input.from_flat(x, self.S, self.has_img)
      ppf.forward(self.P_pf, inp, mean)
      pvf.forward(self.P_vf, inp, value)
    return mean, value

  def act(self, obs, noise=None, row=None):
    """One collector step on the device (reference collector/on_policy.py:90-118 calls
    `pf.explore(ob)` and `vf(ob)` separately, builds and drops two autograd graphs and round-trips twice):
    actions a = mean + std * eps and values for the E observations of one env step, with one shared-encoder pass
    (`infer`).  row = t: the converted observation rows are written straight into the device-resident rollout
    planes at time row t (fp16 space-to-depth image + proprio), so that an epoch whose T rows were all acted on
    here needs no observation H2D in `update_per_epoch` at all.  noise: optional [E, A] standard-normal draw (for
    reproducible tests), else drawn on the device.  Returns host arrays shaped like the reference's outputs."""
    E = int(np.shape(obs)[0]) if not torch.is_tensor(obs) else int(obs.shape[0])
    imgs_out = state_out = None
    r = self._roll
    if row is not None and self.precision == "f16" and self.has_img and r is not None and r["E"] == E and 0 <= row < r["T"]:
      imgs_out = r["imgs"][row * E:(row + 1) * E]
      state_out = r["state"][row * E:(row + 1) * E] if self.S else None
      self._resident_rows = getattr(self, "_resident_rows", set())
      self._resident_rows.add(int(row))
    mean, value = self.infer(obs, imgs_out, state_out)
    log_std = torch.clamp(self.logstd, -5.0, 2.0)
    std = torch.exp(log_std)
    eps = torch.as_tensor(noise, dtype=torch.float32, device=self.device) if noise is not None else torch.randn_like(mean)
    action = mean + std * eps
    ent = (0.5 + 0.5 * np.log(2 * np.pi) + log_std).sum().expand(E, 1)
    out = torch.cat([action, mean, value, ent], 1).cpu().numpy()          # ONE device -> host copy
    A = self.A
    return {"action": out[:, :A], "mean": out[:, A:2 * A], "value": out[:, 2 * A:2 * A + 1], "ent": out[:, 2 * A + 1:],
            "log_std": log_std.cpu().numpy(), "std": std.cpu().numpy()}

  # ---------------------------------------------------------------------------------------------
  # one epoch
  # ---------------------------------------------------------------------------------------------
  def set_lr(self, plr, vlr):
    self.hyper_pf[0] = float(plr)
    self.hyper_vf[0] = float(vlr)

  def sync_target(self):
    """copy_model_params_from_to(pf, target_pf) as one D2D copy (reference utils.py:23-25)."""
    self.t_flat.copy_(self.pf_flat)
    if self.precision == "f16":
      self.plan_t.pack(self.t_flat)

  def _bufs(self, B):
    b = self._mb_bufs.get(B)
    if b is None:
      dev = self.device
      f = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
      b = dict(cur_idx=torch.zeros(B, device=dev, dtype=torch.int32), values=f(B, 1), d_values=f(B, 1),
               mean=f(B, self.A), tmean=f(B, self.A), d_mean=f(B, self.A),
               stats=torch.zeros(8, device=dev, dtype=torch.float64))
      if self.world > 1:
        b["stats_all"] = torch.zeros((self.world, 8), device=dev, dtype=torch.float64)
      if self.precision == "f16":
        b["st"] = torch.zeros((B, self.plan_pf.Sp), device=dev, dtype=torch.float16)
      self._mb_bufs[B] = b
    return b

  def _input(self, B, idx):
    r = self._roll
    if self.has_img:
      return engine.Input(B, r["state"], self.S, 0, r["img"], engine.IMG_ELEMS, 0, idx)
    return engine.Input(B, r["state"], self.S, 0, idx=idx)

  def _minibatch(self, B, with_target=True):
    """The fixed kernel sequence of one PPO.update (reference ppo.py:125-153): critic first,
    then the actor on the encoder the critic just stepped."""
    ops, r, b = self.ops, self._roll, self._bufs(B)
    idx = b["cur_idx"]
    inv_local = 1.0 / B
    inv_global = 1.0 / (B * self.world)
    if self.precision == "f16":
      return self._minibatch_tc(B, b, idx, inv_local, inv_global, with_target)
    ops.select_rows(self._flat_idx, self._slot, idx, B)
    es = self._epoch_stats
    if es is None:
      ops.adv_stats(r["advs"], idx, B, b["stats"])
      if self.world > 1:
        self._allreduce_stats(b)
    inp = self._input(B, idx)
    # ---- critic
    self.plan_vf.forward(self.P_vf, inp, b["values"])
    ops.vf_loss(b["values"], r["rets"], r["values"], idx, b["d_values"], B, inv_global, inv_local,
                self.clipped_value_loss, self.clip_para, self._info, self._slot)
    self.plan_vf.backward(self.P_vf, self.G_vf, b["d_values"])
    if self.world > 1:
      self._allreduce(self.g_vf)
    ops.clip_adam(self.vf_flat, self.g_vf, self.m_vf, self.v_vf, self.n_vf, self.hyper_vf, self._info,
                  self._slot, INFO_GRAD_NORM_VF)
    # ---- actor
    self.plan_pf.forward(self.P_pf, inp, b["mean"])
    self.plan_t.forward(self.P_t, inp, b["tmean"])
    ops.pf_loss(b["mean"], self.logstd, b["tmean"], self.t_logstd, r["acts"], r["advs"], idx,
                b["stats"] if es is None else es, b["d_mean"], self.G_pf["logstd"], B, self.A, inv_global, inv_local,
                self.clip_para, self.entropy_coeff, self._info, self._slot, stats_per_slot=es is not None)
    self.plan_pf.backward(self.P_pf, self.G_pf, b["d_mean"])
    if self.world > 1:
      self._allreduce(self.g_pf)
    ops.clip_adam(self.pf_flat, self.g_pf, self.m_pf, self.v_pf, self.n_pf, self.hyper_pf, self._info,
                  self._slot, INFO_GRAD_NORM_PF)
    ops.slot_advance(self._slot, 0)

  def _minibatch_tc(self, B, b, idx, inv_local, inv_global, with_target=True):
    """Same sequence on the tensor-core tier: fp16 activations, tcgen05 GEMMs.  The chain is kept short:
    one prologue launch, each loss kernel writes the loss-scaled fp16 gradient the backward starts from,
    and ONE optimiser-tail launch per network (split-K reduction of the weight gradients -> clip + Adam ->
    fp16 re-pack of the weights the NEXT forward reads: the critic's tail re-packs the actor's weights,
    whose shared encoder it has just stepped, the actor's tail the critic's)."""
    ops, r = self.ops, self._roll
    imgs, st = r["imgs"], b["st"]
    ppf, pvf = self.plan_pf, self.plan_vf
    es = self._epoch_stats            # advantage statistics of every minibatch, computed once per epoch (run_epoch)
    ops.mb_begin(self._flat_idx, self._slot, idx, B, r["advs"], b["stats"] if es is None else None,
                 r["state"] if self.S else None, self.S, st, ppf.Sp)
    if es is None and self.world > 1:
      self._allreduce_stats(b)
    # The frozen target policy (copied once per update_per_epoch, ppo.py:34) only depends on the rollout
    # row: its action mean is computed when a row is first visited (first opt-epoch), as a parallel
    # branch of the captured graph next to the critic phase, written straight into a per-rollout
    # [N, A] table (row map with the minibatch index list) and re-read by later opt-epochs — the values
    # are bit-identical to recomputing them, 2/3 of the target forwards disappear.
    if with_target:
      with ops.fork(1):
        self.plan_t.forward(self.t_flat, imgs, idx, st, B, r["tmean_all"],
                            out_map=engine.RM(1, self.A, 0, 0, idx=idx))
    # ---- critic (weights were packed by the previous actor tail / run_epoch)
    pvf.forward(self.vf_flat, imgs, idx, st, B, b["values"])
    ops.vf_loss(b["values"], r["rets"], r["values"], idx, b["d_values"], B, inv_global, inv_local,
                self.clipped_value_loss, self.clip_para, self._info, self._slot,
                d_f16=pvf.grad_in(B), scale_f16=pvf.loss_scale(B))
    pvf.backward(self.g_vf, None, flush=False)
    self._tail("vf", None)
    # ---- actor
    ppf.forward(self.pf_flat, imgs, idx, st, B, b["mean"])
    if with_target:
      ops.join(1)
    ops.pf_loss(b["mean"], self.logstd, r["tmean_all"], self.t_logstd, r["acts"], r["advs"], idx,
                b["stats"] if es is None else es, b["d_mean"], self.G_pf["logstd"], B, self.A, inv_global, inv_local,
                self.clip_para, self.entropy_coeff, self._info, self._slot, target_indexed=True,
                d_f16=ppf.grad_in(B), scale_f16=ppf.loss_scale(B), stats_per_slot=es is not None)
    ppf.backward(self.g_pf, None, flush=False)
    self._tail("pf", self._slot)

  def _tail(self, which, slot_advance):
    """reduce -> (all-reduce) -> clip + Adam (+ fp16 operand copies of both networks) -> counters"""
    if which == "vf":
      flat, g, m, v, n, hyper, norm_slot = self.vf_flat, self.g_vf, self.m_vf, self.v_vf, self.n_vf, self.hyper_vf, INFO_GRAD_NORM_VF
      W_self, W_other, extra = self.plan_vf.W, self.plan_pf.W, (0, 0)
    else:
      flat, g, m, v, n, hyper, norm_slot = self.pf_flat, self.g_pf, self.m_pf, self.v_pf, self.n_pf, self.hyper_pf, INFO_GRAD_NORM_PF
      W_self, W_other = self.plan_pf.W, self.plan_vf.W
      off = self.pf_layout["logstd"][0]
      extra = (off, self.A)
    kw = dict(param=flat, grad=g, m=m, v=v, n=n, hyper=hyper, info=self._info, slot=self._slot,
              norm_slot=norm_slot, extra=extra, scatter=self.scatter[which], packed_self=W_self.packed,
              packed_other=W_other.packed, slot_advance=slot_advance)
    if self.world > 1:
      self.ops.opt_tail(1)
      self._allreduce(g)
      self.ops.opt_tail(2, **kw)
    else:
      self.ops.opt_tail(3, **kw)

  def _allreduce(self, t):
    import torch.distributed as dist
    dist.all_reduce(t, group=self.pg)

  def _allreduce_stats(self, b):
    """global (sum, sumsq, n, max, min) of the advantages across ranks: one all-gather of 8
    doubles, combined on the device (advantage normalisation is over the GLOBAL minibatch)."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(b["stats_all"].view(-1), b["stats"], group=self.pg)
    a = b["stats_all"]
    b["stats"][0:3] = a[:, 0:3].sum(0)
    b["stats"][3] = a[:, 3].max()
    b["stats"][4] = a[:, 4].min()

  def run_epoch(self, perms, batch_size):
    """opt_epochs x minibatches.  perms: [opt_epochs, T] time-row permutations.  Returns the
    per-minibatch info dicts (one D2H read of the info table)."""
    self.check_views()
    r = self._roll
    T, E = r["T"], r["E"]
    assert batch_size % E == 0, "batch size should be dividable by env_nums"
    rows = batch_size // E
    perms = np.asarray(perms)
    flat = (perms[:, :, None].astype(np.int64) * E + np.arange(E)[None, None, :]).reshape(len(perms), T * E)
    n_full, tail = divmod(T, rows)
    n_mb = len(perms) * (n_full + (1 if tail else 0))
    dev = self.device
    self._flat_idx = torch.from_numpy(flat.astype(np.int32)).to(dev)
    self._slot = getattr(self, "_slot", None)
    if self._slot is None:
      self._slot = torch.zeros(1, device=dev, dtype=torch.int32)
    if getattr(self, "_info", None) is None or self._info.shape[0] < n_mb:
      self._info = torch.zeros((n_mb, INFO_STRIDE), device=dev, dtype=torch.float32)
      self._graphs.clear()
    B = rows * E
    if self.precision == "f16":
      # parameters may have been changed from outside (load_state_dict) since the last epoch: re-pack both
      # operand copies once; inside the epoch the optimiser tails keep them current (_minibatch_tc)
      self.plan_vf.pack(self.vf_flat)
      self.plan_pf.pack(self.pf_flat)
    if tail == 0 and len(perms) > 0:
      # uniform minibatches: flat_idx is [n_mb, B] and the device slot counter indexes it
      if getattr(self, "_flat_idx_static", None) is None or self._flat_idx_static.numel() != flat.size:
        self._flat_idx_static = torch.empty(flat.size, device=dev, dtype=torch.int32)
        self._graphs.clear()
      self._flat_idx_static.copy_(self._flat_idx.view(-1))
      self._flat_idx = self._flat_idx_static
      self._slot.zero_()
      # advantage statistics of all minibatches in ONE launch (and, data parallel, ONE exchange per epoch instead
      # of one per minibatch on the critical path): rows and advantages are fixed from here on
      st_all = getattr(self, "_stats_table", None)
      if st_all is None or st_all.shape[0] != n_mb:
        st_all = self._stats_table = torch.zeros((n_mb, 8), device=dev, dtype=torch.float64)
        self._graphs.clear()
      self.ops.adv_stats_epoch(self._flat_idx, n_mb, B, r["advs"], st_all)
      if self.world > 1:
        import torch.distributed as dist
        allst = torch.empty((self.world, n_mb, 8), device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allst.view(-1), st_all.view(-1), group=self.pg)
        st_all[:, 0:3] = allst[:, :, 0:3].sum(0)
        st_all[:, 3] = allst[:, :, 3].max(0).values
        st_all[:, 4] = allst[:, :, 4].min(0).values
      self._epoch_stats = st_all
      pending, self._pending_obs = getattr(self, "_pending_obs", None), None
      self._pending_half_cur, self._pending_half = getattr(self, "_pending_half", None), None
      cur = torch.cuda.current_stream(dev)
      if pending is not None:
        with self.ops.fork(2):          # order the copy stream after the index upload / previous epoch
          pass
      k = 0
      while k < n_mb:
        if pending is not None and k < n_full:
          # interleaved with the launches so the CPU never runs far behind the GPU
          ev = self._stream_chunk(pending[0], pending[1], perms[0][k * rows:(k + 1) * rows], E, k, rows)
          cur.wait_event(ev)               # rows of minibatch k (first opt-epoch) have landed
        # later opt-epochs need no per-minibatch gating: GRAPH_GROUP minibatches per graph replay (one launch
        # latency per group instead of per minibatch)
        group = GRAPH_GROUP if (k >= n_full and n_mb - k >= GRAPH_GROUP and self.precision == "f16") else 1
        self._launch(B, with_target=k < n_full, count=group)
        k += group
    else:
      if getattr(self, "_pending_obs", None) is not None:
        obs, D = self._pending_obs
        self._pending_obs = None
        if self._pending_half is not None:    # ragged epochs copy the fp32 rows up front
          self.h2d_bytes += T * E * D * 4 - T * E * (engine.IMG_ELEMS * 2 + self.S * 4)
          self._pending_half = None
        self._copy_obs_rows(obs, D, 0, T * E)
      self._run_ragged(flat, T, E, rows)
    self._epoch_stats = None
    info = self._info[:n_mb, :len(INFO_KEYS)].cpu().numpy()
    self.d2h_bytes = info.nbytes
    return [dict(zip(INFO_KEYS, (float(x) for x in row))) for row in info]

  def _launch(self, B, with_target=True, count=1):
    """with_target: this minibatch visits its rows for the first time in this epoch (first opt-epoch):
    the graph variant that also runs the frozen target policy's forward.  count: consecutive minibatches
    captured in (and replayed as) one graph."""
    if self.precision != "f16":
      with_target = True                      # the exact tier recomputes the target forward every time
    if not self.use_cuda_graph:
      for _ in range(count):
        self._minibatch(B, with_target)
      return
    key = (B, with_target, count)
    g = self._graphs.get(key)
    if g is None:
      if not self._graphs.get(("warm", B)):
        # first minibatch at this size runs eagerly: allocates workspaces, loads modules
        for _ in range(count):
          self._minibatch(B, with_target)
        self._graphs[("warm", B)] = True
        return
      torch.cuda.synchronize(self.device)
      g = torch.cuda.CUDAGraph()
      launches0 = self.ops.launches
      with torch.cuda.graph(g):
        for _ in range(count):
          self._minibatch(B, with_target)
      self._graphs[key] = (g, self.ops.launches - launches0)
      g = self._graphs[key]
    else:
      self.ops.launches += g[1]
    g[0].replay()

  def _run_ragged(self, flat, T, E, rows):
    """T not divisible by the minibatch rows: the reference yields a short last minibatch."""
    slot = 0
    saved = self._flat_idx
    for ep in range(flat.shape[0]):
      for pos in range(0, T, rows):
        n = min(rows, T - pos) * E
        self._flat_idx = saved[ep, pos * E:pos * E + n].contiguous()
        self._slot.fill_(0)
        info_row = self._info
        self._info = info_row[slot:slot + 1]
        self._minibatch(n)
        self._info = info_row
        slot += 1
    self._flat_idx = saved
