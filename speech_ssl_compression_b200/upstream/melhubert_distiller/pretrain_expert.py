"""Reference ``upstream/melhubert_distiller/pretrain_expert.py`` differs from
``distillation/pretrain_expert.py`` only in reading the student's config from the key
``student`` instead of ``melhubert`` (lines 13/46/61); the shipped yamls use ``melhubert``
(SURVEY Q2), so this variant accepts either."""
from ...distillation.pretrain_expert import MelHuBERTDistiller as _Base


class MelHuBERTDistiller(_Base):
    def __init__(self, upstream_config, initial_weight=None, device="cuda", multi_gpu=False):
        if "student" in upstream_config and "melhubert" not in upstream_config:
            self.STUDENT_KEY = "student"
        super().__init__(upstream_config, initial_weight, device, multi_gpu)
