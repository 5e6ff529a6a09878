"""Multi-generator scheduling policy (host Python; negligible time).

The reference ships this design only as README prose (readme.md:2-10, translated in SURVEY Appendix B) plus two helper
functions (src/utils.py:102-110 ``shuffle_lists_in_same_order``, src/utils.py:113-115 ``interpolate_models``) and a
stale call that hints at three generators (src/main.py:28).  There is no executable oracle, so the rules below ARE the
frozen specification the tests pin (tests/test_policy.py); every constant the README leaves open is a named field.

README rules implemented (line numbers in readme.md):
  (4) keep the generators sorted by contrast ("com") loss, ascending; every batch trains them in that order;
  (5,6) a generator compares its own contrast loss with ``pre_loss``: larger -> prefer the contrast loss, otherwise
        use the discriminator;
  (7) the first generator is the main producer of new information;
  (8) re-sort after every epoch;
  (10) two regimes split at ``starting_gan_loss``: above it -> contrast loss with high probability, GAN with low
       probability; below it -> the first generator uses GAN with high probability, a later generator whose contrast
       loss exceeds the current minimum (``pre_loss``) uses the contrast loss with high probability;
  (13) optional "strong leads weak": p <- alpha * p_best + (1 - alpha) * p  (interpolate_models, alpha 0.2).
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

PIXEL = "pixel"
GAN = "gan"


def shuffle_lists_in_same_order(*lists):
    """Same contract as src/utils.py:102-110: zip the lists, sort DESCENDING by the last list (stable), unzip."""
    combined = list(zip(*lists))
    combined.sort(key=lambda t: t[-1], reverse=True)
    return [list(t) for t in zip(*combined)]


@dataclass
class PolicyConfig:
    num_generators: int = 3            # inferred from `train_example(20, 3)` (src/main.py:28)
    starting_gan_loss: float = 0.05    # README "Starting_GAN_loss" (value unspecified upstream)
    p_high: float = 0.9                # README "high probability"
    p_low: float = 0.1                 # README "low probability"
    force: Optional[str] = None        # PIXEL: pixel-loss pre-training phase; GAN: fine-tune phase (BASELINE cfg5)
    seed: int = 0                      # shared by all ranks so decisions agree (random is imported at src/train.py:2)
    lead_alpha: float = 0.0            # >0 enables "strong leads weak" at epoch end (src/utils.py:113-115 uses 0.2)


def gan_probability(position: int, own_loss: float, pre_loss: float, cfg: PolicyConfig) -> float:
    """Probability that the generator at ``position`` in the sorted list trains against the discriminator.

    ``own_loss``: its running contrast loss; ``pre_loss``: the loss it is compared with (README (5)): the running
    contrast loss of the generator ranked just before it, i.e. the current minimum for position 1."""
    if cfg.force == PIXEL:
        return 0.0
    if cfg.force == GAN:
        return 1.0
    if not (own_loss <= cfg.starting_gan_loss):          # regime 1 (also taken while the loss is still NaN/unknown)
        return cfg.p_low
    if position == 0:                                     # regime 2, leader
        return cfg.p_high
    return cfg.p_low if own_loss > pre_loss else cfg.p_high


def decide(position: int, own_loss: float, pre_loss: float, cfg: PolicyConfig, u: float) -> str:
    """The frozen decision function: ``u`` ~ U[0,1) from the shared seeded RNG."""
    return GAN if u < gan_probability(position, own_loss, pre_loss, cfg) else PIXEL


@dataclass
class MultiGeneratorPolicy:
    """Tracks running contrast losses, produces the per-batch training order and PIXEL/GAN decisions."""
    cfg: PolicyConfig = field(default_factory=PolicyConfig)

    def __post_init__(self):
        k = self.cfg.num_generators
        self.rng = random.Random(self.cfg.seed)
        self.order: List[int] = list(range(k))            # generator ids, best (lowest contrast loss) first
        self.epoch_sum = [0.0] * k
        self.epoch_cnt = [0] * k
        self.running = [float("inf")] * k                 # last known mean contrast loss per generator id

    # ---- per batch --------------------------------------------------------------------------------------------
    def plan_batch(self) -> List[tuple]:
        """[(generator id, PIXEL | GAN), ...] in training order for the next batch.  Exactly one RNG draw per
        generator per batch, whatever the outcome, so all ranks stay in lock-step."""
        plan = []
        for pos, gid in enumerate(self.order):
            own = self.running[gid]
            pre = self.running[self.order[pos - 1]] if pos > 0 else own
            plan.append((gid, decide(pos, own, pre, self.cfg, self.rng.random())))
        return plan

    def observe(self, gid: int, com_loss: float) -> None:
        """Feed the contrast loss a generator just obtained (mean over ranks under data parallelism)."""
        self.epoch_sum[gid] += float(com_loss)
        self.epoch_cnt[gid] += 1
        self.running[gid] = self.epoch_sum[gid] / self.epoch_cnt[gid]

    # ---- per epoch --------------------------------------------------------------------------------------------
    def end_epoch(self) -> List[int]:
        """README (8): re-sort ascending by epoch-mean contrast loss (ties keep the previous order), reset sums."""
        ids = list(self.order)
        scores = [-self.running[g] for g in ids]          # shuffle_lists_in_same_order sorts descending
        ids, _ = shuffle_lists_in_same_order(ids, scores)
        self.order = ids
        k = self.cfg.num_generators
        self.epoch_sum = [0.0] * k
        self.epoch_cnt = [0] * k
        return list(self.order)


def interpolate_models(model, target_model, alpha: float = 0.2) -> None:
    """"Strong leads weak" (src/utils.py:113-115): param <- alpha * target + (1 - alpha) * param, parameters only.
    In place on the flat buffers so the nn.Parameter views stay valid."""
    import torch
    with torch.no_grad():
        flat, tflat = model.flat_parameters(), target_model.flat_parameters()
        flat.mul_(1.0 - alpha).add_(tflat, alpha=alpha)
