"""Ensemble of independent solves (BASELINE configs[4], SURVEY 8e "Ensemble config"): many members
of the same solver classes - different initialisations (seed) and frequency scales - advanced
together.  The reference runs such sweeps as one process per member (run_1d.sh / run_2d.sh loop
over equations and kernels; model_GP_solver_2d.py:235-352 is the per-member loop); here the
members share a GPU:

  * every member owns its plan (gphm_plan) and packed state, so members never touch each other;
  * one ensemble step = one `gphm_step` per member, issued round-robin on a few CUDA streams so
    that the small kernels of different members overlap (an N = 400 problem fills a fraction of
    the 148 SMs), and recorded ONCE into a CUDA graph: a step of the whole ensemble is then a
    single graph launch instead of ~40 kernel launches per member (the launch-bound regime);
  * across GPUs the members are partitioned (`shard_members`), one process per GPU, NO data-path
    collective; the per-member results (loss, rel-L2 error) are gathered at the end
    (`gather_results`, works over NCCL and gloo).

Initialisation of member (seed s, freq_scale fs), SURVEY 8d: the reference's init
(model_GP_solver_2d.py:245-261) with `freq = linspace(0,1,Q) * fs + 0.01 * N(0,1)` drawn from
`torch.Generator().manual_seed(s)` (the reference itself is deterministic: its PRNGKey is unused).
"""
import contextlib
import io

import numpy as np
import torch

DEFAULT_FREQ_SCALES = (10, 20, 30, 40, 50, 60, 70, 80)


def ensemble_members(n_seeds=64, freq_scales=DEFAULT_FREQ_SCALES):
    """[(seed, freq_scale)] - 64 x 8 = 512 members by default, seed-major."""
    return [(s, fs) for s in range(int(n_seeds)) for fs in freq_scales]


def shard_members(n_members, rank, world):
    """Contiguous block of member indices owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside [0, %d)" % (rank, world))
    base, extra = divmod(int(n_members), int(world))
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def member_init(model, seed, freq_scale):
    """params pytree of one member: model.init_params() with the member's frequency initialisation."""
    params = model.init_params()
    gen = torch.Generator().manual_seed(int(seed))
    for key in ("kernel_paras", "kernel_paras_1", "kernel_paras_2"):
        if key in params:
            Q = len(params[key]["freq"])
            noise = 0.01 * torch.randn(Q, generator=gen, dtype=torch.float64).numpy()
            params[key]["freq"] = np.linspace(0, 1, Q) * float(freq_scale) + noise
    return params


def gather_results(local, n_members, rank, world, group=None):
    """Concatenate the per-rank result rows (n_local, k) in member order on every rank."""
    if local.dim() != 2:
        raise ValueError("gather_results: (n_local, k) rows expected")       # an empty shard still knows k
    if world == 1:
        return local.clone()
    import torch.distributed as dist
    k = local.shape[1]
    cap = max(len(shard_members(n_members, r, world)) for r in range(world))
    pad = torch.zeros((cap, k), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][:len(shard_members(n_members, r, world))] for r in range(world)], dim=0)


class Ensemble(object):
    """Members = solver objects of this package (GP_solver_1d_single, GP_solver_2d_single,
    GP_solver_2d_single_advection: anything with `.core`, `.lr`, `.init_params()`), one per member,
    with their initial params.  `step()` advances every member by one Adam iteration."""

    def __init__(self, models, params_list, streams=16, graph=True):
        if len(models) != len(params_list) or not models:
            raise ValueError("need one params pytree per member")
        self.models = list(models)
        self.states = [m.core.new_state(p) for m, p in zip(self.models, params_list)]
        self.device = self.models[0].core.device
        self.n_streams = max(1, min(int(streams), len(self.models)))
        self.side = [torch.cuda.Stream(device=self.device) for _ in range(self.n_streams)]
        self.use_graph = bool(graph)
        self.graph = None
        self.steps_done = 0
        self.kernels_per_step = None       # libgphm kernel launches one ensemble step issues (counted at capture / first issue)

    def __len__(self):
        return len(self.models)

    def issue_mode(self):
        return ("CUDA-graph replay, %d streams" if self.use_graph else "plain launches, %d streams") % self.n_streams

    def _issue(self):
        """One gphm_step per member, round-robin over the side streams, joined on the current stream."""
        cur = torch.cuda.current_stream(self.device)
        lib = self.models[0].core.lib
        c0 = lib.gphm_launch_count()
        for s in self.side:
            s.wait_stream(cur)
        for i, (m, st) in enumerate(zip(self.models, self.states)):
            with torch.cuda.stream(self.side[i % self.n_streams]):
                m.core.step_inplace(st, m.lr)
        for s in self.side:
            cur.wait_stream(s)
        self.kernels_per_step = int(lib.gphm_launch_count() - c0)

    def _capture(self):
        # lazy one-time initialisation inside libgphm (function attributes) must not happen under capture:
        # run one step of member 0's plan on a scratch copy of its state first
        m0 = self.models[0]
        scratch = m0.core.new_state(m0.core.unpack_tree(self.states[0].U, self.states[0].small))
        m0.core.step_inplace(scratch, m0.lr)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._issue()
        self.graph = g

    def step(self, n=1):
        """n Adam iterations of every member (no host synchronisation)."""
        for _ in range(int(n)):
            if self.use_graph:
                if self.graph is None:
                    self._capture()
                self.graph.replay()
            else:
                self._issue()
            self.steps_done += 1

    def losses(self):
        """(B,) device tensor: the loss each member's LAST step evaluated (before its update)."""
        return torch.stack([st.terms[0] for st in self.states])

    def errors(self):
        """(B,) device tensor: rel-L2 error of every member's prediction on its test grid."""
        out = []
        for m, st in zip(self.models, self.states):
            if m.core.dim == 2:
                pred = m.core.predict(st, m.Xte[0], m.Xte[1])
                out.append(m.core.rel_l2(pred, m.ute))
            else:
                out.append(m.core.rel_l2(m.core.predict(st, m.Xte), m.yte))
        return torch.cat(out)

    def params(self, i):
        return self.models[i].core.unpack_tree(self.states[i].U, self.states[i].small)

    def raise_on_bad_status(self):
        for m in self.models:
            m.core.raise_on_bad_status()


def build_ensemble(trick_paras, members, rank=0, world=1, streams=16, graph=True, quiet=True, batched=False):
    """Ensemble of this rank's share of `members` [(seed, freq_scale)] for one equation config
    (`trick_paras` as evals() builds it; its `equation` prefix selects the solver class).
    Returns (Ensemble, indices of the members it holds)."""
    from . import model_GP_solver_1d as m1d, model_GP_solver_2d as m2d, model_GP_solver_advection as madv
    eq_type = trick_paras["equation"].split("-")[0]
    mine = shard_members(len(members), rank, world)
    sink = io.StringIO()
    with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
        if eq_type in ("poisson_1d", "allencahn_1d"):
            Xind, y, X_col, src, X_test, Y_test = m1d.build_problem(trick_paras)
            make = lambda tp: m1d.GP_solver_1d_single(Xind, y, X_col, src, 1e-6, X_test, Y_test, tp)
        elif eq_type == "advection":
            prob = madv.build_problem(trick_paras)
            make = lambda tp: madv.GP_solver_2d_single_advection(*prob[:3], 1e-6, *prob[3:], tp)
        else:
            prob = m2d.build_problem(trick_paras)
            make = lambda tp: m2d.GP_solver_2d_single(*prob[:3], 1e-6, *prob[3:], tp)
        models, inits = [], []
        for i in mine:
            seed, fs = members[i]
            tp = dict(trick_paras, freq_scale=fs)
            model = make(tp)
            models.append(model)
            inits.append(member_init(model, seed, fs))
    return Ensemble(models, inits, streams=streams, graph=graph), list(mine)
