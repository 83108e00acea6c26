"""Sampler pre-screening: the reference's samplers draw their random
states in buckets of 10 000 (``RandomnessManager.redraw``, randomness.py:150-175), hand them
out one at a time from the END of the bucket (``get_random``, :177-190) and test each with
``cc.feasible`` (``Sampler.get_valid_next_pos``, samplers/baseSampler.py:58-66;
``RandomPolicySampler.get_next_pos``, samplers/randomPolicySampler.py:60-69).

:func:`attach_prescreen` hooks the UNMODIFIED reference objects: every time a bucket is
(re)drawn the whole bucket goes through ``engine.transform`` and one ``feasible_batch``
launch (``cc.prescreen``), and the per-state ``cc.feasible`` calls that follow are answered
from the remembered booleans.  The random stream, the order of the draws, the accepted
states and the ``Stats`` counters are unchanged; only the number of device launches is.
"""
from __future__ import annotations


def attach_prescreen(randomness_manager, engine):
    """Wrap ``randomness_manager.redraw`` (a reference ``RandomnessManager``, or anything
    with ``redraw(method)`` and ``random_draws[method]``) so that each fresh bucket is
    pre-screened by ``engine.cc``.  Returns the manager."""
    if getattr(randomness_manager, "_sbp_prescreen", False):
        return randomness_manager
    original = randomness_manager.redraw

    def redraw(random_method):
        original(random_method)
        bucket = randomness_manager.random_draws[random_method]
        # the same transform the sampler applies to each draw (engine.py:50-57): elementwise,
        # so a row of the transformed bucket is bit-identical to transform(row)
        engine.cc.prescreen(engine.transform(bucket))

    randomness_manager.redraw = redraw
    randomness_manager._sbp_prescreen = True
    return randomness_manager


def attach_prescreen_to_sampler(sampler, engine=None):
    """Convenience for the reference's ``RandomPolicySampler`` family: ``sampler.random`` is
    its ``RandomnessManager`` (created by ``sampler.init``), ``sampler.args.engine`` the engine."""
    engine = engine if engine is not None else sampler.args.engine
    attach_prescreen(sampler.random, engine)
    return sampler
