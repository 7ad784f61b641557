from .pretrain_expert import MelHuBERTPretrainer as UpstreamPretrainExpert  # noqa: F401
