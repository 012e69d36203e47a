"""A minimal stand-in for the reference's global `hparams` object (hyperparams.py:3-72) carrying
the same names and default values, for callers that do not have the reference checkout on the
path (tests, bench.py).  When the reference's own train.py / eval.py drive this package they pass
their own HParams object; the modules only read attributes."""
import copy


class HParams:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def set_hparam(self, name, value):
        if name not in self.__dict__:
            raise KeyError(name)
        setattr(self, name, value)

    def parse(self, spec):
        """"a=1,b=0.5,c=True" -> typed by the current value, like utils/hparams.py:401-418."""
        for item in filter(None, (s.strip() for s in spec.split(","))):
            k, v = item.split("=", 1)
            cur = self.__dict__[k]
            if isinstance(cur, bool):
                v = v.lower() in ("1", "true", "yes")
            else:
                v = type(cur)(v)
            setattr(self, k, v)
        return self

    def copy(self, **overrides):
        out = copy.deepcopy(self)
        for k, v in overrides.items():
            out.set_hparam(k, v)
        return out

    def values(self):
        return dict(self.__dict__)


def default_hparams():
    return HParams(
        num_mels=80, max_generation_frames=1100, vocab_size=6000, embed_size=512, encoder_hidden=512,
        decoder_hidden=768, n_encoder_layer=6, n_decoder_layer=6, n_attention_head=8,
        transformer_dropout_rate=0.1, decoder_dropout_rate=0.5, prenet_hidden=256, postnet_hidden=512,
        n_postnet_layer=5, reg_weight=5e-9, multi_speaker=True, max_num_speaker=1000,
        speaker_embedding_size=128, multi_lingual=True, max_num_language=100, language_net_hidden=128,
        language_embedding_size=128, warmup_steps=50000, max_lr=1e-3, min_lr=1e-5, lr_decay_step=550000,
        lr_decay_rate=1e-2, adam_eps=5e-8)


def hparams_from(cfg):
    """HParams with the model-shaping values of an oracle-style config object."""
    hp = default_hparams()
    for k, v in vars(cfg).items():
        setattr(hp, k, v)
    return hp
