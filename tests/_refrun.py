"""Shared harness that drives the UNMODIFIED reference planners (imported through
oracle/ref_shims.py: from /root/reference in the build container, from the staged copy
oracle/_ref on the GPU box) for whole seeded runs.  Test infrastructure only."""
from __future__ import annotations

import numpy as np

import ref_shims


def reference_available() -> bool:
    return ref_shims.reference_available()


def silence_tqdm(monkeypatch):
    from functools import partialmethod

    from tqdm import tqdm

    monkeypatch.setattr(tqdm, "__init__", partialmethod(tqdm.__init__, disable=True))


def make_args(planner_id, engine_factory, start, goal, max_nodes, **kw):
    """PlanningOptions built the way the reference's tests/common_vars.py:44-81 does
    (generate_args needs docopt, which is stubbed)."""
    from sbp_env.utils import planner_registry
    from sbp_env.utils.common import PlanningOptions
    import sbp_env.planners  # noqa: F401 - fills the registries
    import sbp_env.samplers  # noqa: F401

    pdp = planner_registry.PLANNERS[planner_id]
    opts = dict(always_refresh=False, epsilon=None, goalBias=0.03, start_pt=start, goal_pt=goal,
                goal_radius=None, ignore_step_size=False, max_number_nodes=max_nodes, no_display=True,
                planner_data_pack=pdp, sampler_data_pack=planner_registry.SAMPLERS[pdp.sampler_id],
                radius=None, rrdt_proposal_distribution="dynamic-vonmises", scaling=1.5,
                showSampledPoint=False, skip_optimality=False)
    opts.update(kw)
    args = PlanningOptions(**opts)
    args.engine = engine_factory(args)
    return args


def run(planner_id, engine_factory, swap_cc, start, goal, max_nodes, before_run=None, **kw):
    """One seeded run (fixed_seed=0).  ``swap_cc(cc)`` replaces the engine's checker before
    Env is built; ``before_run(env)`` runs after Env.__init__ (hooks go here)."""
    from sbp_env.env import Env
    from sbp_env.utils.common import Stats

    Stats.clear_instance()
    args = make_args(planner_id, engine_factory, start, goal, max_nodes, **kw)
    if swap_cc is not None:
        args.engine.cc = swap_cc(args.engine.cc)
    env = Env(args, fixed_seed=0)
    if before_run is not None:
        before_run(env)
    stats = env.run()
    pl = env.planner
    n = len(pl.nodes)
    return dict(poses=np.array(pl.poses[:n]), c_max=pl.c_max, n=n,
                counters=(stats.visible_cnt, stats.feasible_cnt, stats.valid_sample,
                          stats.invalid_samples_obstacles, stats.invalid_samples_connections),
                env=env)


def same_run(a, b):
    assert a["n"] == b["n"]
    assert np.array_equal(a["poses"], b["poses"])
    assert a["c_max"] == b["c_max"]
    assert a["counters"] == b["counters"]


def tree_state(planner, nodes=None):
    """(costs, parent positions) of a planner's node list, for whole-tree comparison."""
    nodes = planner.nodes if nodes is None else nodes
    costs = [float(n.cost) for n in nodes]
    parents = [None if n.parent is None else tuple(np.asarray(n.parent.pos).tolist()) for n in nodes]
    return costs, parents
