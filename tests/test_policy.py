"""Table-driven tests of the multi-generator policy (the frozen spec of SURVEY Appendix B / readme.md:2-10)."""
import random

import pytest

import srgan_b200 as S
from srgan_b200 import GAN, PIXEL, MultiGeneratorPolicy, PolicyConfig, decide, gan_probability, shuffle_lists_in_same_order


def test_shuffle_lists_in_same_order_contract():
    # src/utils.py:102-110: sort descending by the LAST list, stable, all lists permuted alike
    a, b, c = ["g0", "g1", "g2", "g3"], [10, 11, 12, 13], [0.5, 0.9, 0.5, 0.1]
    assert shuffle_lists_in_same_order(a, b, c) == [["g1", "g0", "g2", "g3"], [11, 10, 12, 13], [0.9, 0.5, 0.5, 0.1]]


CFG = PolicyConfig(num_generators=3, starting_gan_loss=0.05, p_high=0.9, p_low=0.1)


@pytest.mark.parametrize("pos,own,pre,expected", [
    (0, 0.30, 0.30, 0.1),      # regime 1: loss above Starting_GAN_loss -> GAN with low probability
    (2, 0.30, 0.20, 0.1),
    (0, 0.04, 0.04, 0.9),      # regime 2: the leader uses GAN with high probability
    (1, 0.045, 0.04, 0.1),     # regime 2, later model, loss above the loss it is compared with -> contrast loss
    (1, 0.03, 0.04, 0.9),      # regime 2, later model, not worse -> GAN
    (1, 0.04, 0.04, 0.9),      # tie is "not larger"
    (0, float("inf"), float("inf"), 0.1),   # no loss observed yet
    (0, 0.05, 0.05, 0.9),      # boundary: loss == Starting_GAN_loss counts as reached
])
def test_gan_probability_table(pos, own, pre, expected):
    assert gan_probability(pos, own, pre, CFG) == expected


def test_decide_uses_u_below_probability():
    assert decide(0, 0.04, 0.04, CFG, 0.89) == GAN
    assert decide(0, 0.04, 0.04, CFG, 0.90) == PIXEL
    assert decide(0, 0.30, 0.30, CFG, 0.05) == GAN
    assert decide(0, 0.30, 0.30, CFG, 0.10) == PIXEL


def test_forced_phases():
    pix = PolicyConfig(num_generators=3, force=PIXEL)
    gan = PolicyConfig(num_generators=3, force=GAN)     # BASELINE cfg5: all generators in discriminator mode
    for pos in range(3):
        assert gan_probability(pos, 0.01, 0.5, pix) == 0.0
        assert gan_probability(pos, 0.9, 0.1, gan) == 1.0


def test_policy_sequence_reproducible_and_resort():
    """Feeding the same loss sequence gives the same decisions; the order is re-sorted ascending at epoch end."""
    losses = {0: [0.30, 0.20, 0.04, 0.03], 1: [0.10, 0.04, 0.03, 0.02], 2: [0.50, 0.40, 0.30, 0.06]}

    def run():
        pol = MultiGeneratorPolicy(PolicyConfig(num_generators=3, seed=7))
        plans = []
        for t in range(4):
            plan = pol.plan_batch()
            plans.append(plan)
            for gid, _ in plan:
                pol.observe(gid, losses[gid][t])
        order = pol.end_epoch()
        return plans, order, pol

    p1, o1, pol = run()
    p2, o2, _ = run()
    assert p1 == p2 and o1 == o2
    assert [gid for gid, _ in p1[0]] == [0, 1, 2]              # initial order = construction order
    assert o1 == [1, 0, 2]                                      # ascending epoch-mean contrast loss
    assert [gid for gid, _ in pol.plan_batch()] == [1, 0, 2]
    # exactly one RNG draw per generator per batch, whatever the decision
    rng = random.Random(7)
    us = [rng.random() for _ in range(12)]
    pol2 = MultiGeneratorPolicy(PolicyConfig(num_generators=3, seed=7))
    first = pol2.plan_batch()
    assert [m for _, m in first] == [GAN if u < 0.1 else PIXEL for u in us[:3]]    # all losses unknown -> p_low


def test_policy_rejects_mismatched_generator_count():
    with pytest.raises(ValueError):
        S.MultiGeneratorGAN([object(), object()], [None, None], None, policy=MultiGeneratorPolicy(PolicyConfig(num_generators=3)))
