"""Hyper-parameter holders of the newsrec models on the B200 hot path (NRMS, NRMSDocVec, NAML).

API surface of the reference's src/ebrec/models/newsrec/model_config.py:1-114: plain classes named
``hparams_<model>`` whose class attributes carry the defaults, are read as ``hparams.<name>`` by the models
and are mutated in place by the scripts (examples/reproducibility_scripts/ebnerd_nrms.py:78-96), plus
``print_hparams`` / ``hparams_to_dict`` which walk ``__annotations__``.  Here the holders are generated from
one table of (name, type, default) rows per model, so the shared optimizer block is stated once.
"""
from __future__ import annotations

DEFAULT_TITLE_SIZE = 30          # tokens per title (xlm-roberta ids)
DEFAULT_BODY_SIZE = 40           # tokens per body (NAML)
DEFAULT_DOCUMENT_SIZE = 768      # width of a precomputed document vector (NRMSDocVec)
UNKNOWN_TITLE_VALUE = [0] * DEFAULT_TITLE_SIZE
UNKNOWN_BODY_VALUE = [0] * DEFAULT_BODY_SIZE

# every model trains with the same optimizer block (model_config.py:38-42, 92-95, 109-112)
_TRAINING = (("optimizer", str, "adam"), ("loss", str, "cross_entropy_loss"), ("dropout", float, 0.2),
             ("learning_rate", float, 1e-4))
_DENSE_STACK = "newsencoder_units_per_layer"


def _holder(name: str, *groups) -> type:
    """A class whose attributes are the defaults and whose __annotations__ lists them in declaration order."""
    fields = [row for g in groups for row in g]
    ns = {fname: default for fname, _, default in fields}
    ns["__annotations__"] = {fname: ftype for fname, ftype, _ in fields}
    ns["__doc__"] = f"Default hyper-parameters of {name[len('hparams_'):]} (mutable class attributes)."
    return type(name, (), ns)


hparams_nrms = _holder(
    "hparams_nrms",
    (("title_size", int, DEFAULT_TITLE_SIZE), ("history_size", int, 20)),
    (("head_num", int, 20), ("head_dim", int, 20), ("attention_hidden_dim", int, 200)),
    _TRAINING,
    # optional Dense/BatchNorm/Dropout stack between self-attention and pooling (nrms.py:142-152)
    ((_DENSE_STACK, "list[int]", None), ("newsencoder_l2_regularization", float, 1e-4)),
)

hparams_nrms_docvec = _holder(
    "hparams_nrms_docvec",
    (("title_size", int, DEFAULT_DOCUMENT_SIZE), ("history_size", int, 20)),
    (("head_num", int, 16), ("head_dim", int, 16), ("attention_hidden_dim", int, 200)),
    _TRAINING,
    ((_DENSE_STACK, "list[int]", [512, 512, 512]), ("newsencoder_l2_regularization", float, 1e-4)),
)

hparams_naml = _holder(
    "hparams_naml",
    (("title_size", int, DEFAULT_TITLE_SIZE), ("history_size", int, 20), ("body_size", int, DEFAULT_BODY_SIZE),
     ("vert_num", int, 100), ("vert_emb_dim", int, 10), ("subvert_num", int, 100), ("subvert_emb_dim", int, 10)),
    (("dense_activation", str, "relu"), ("cnn_activation", str, "relu"), ("attention_hidden_dim", int, 200),
     ("filter_num", int, 400), ("window_size", int, 3)),
    _TRAINING,
)


def hparams_to_dict(hparams_class) -> dict:
    """{attribute: current value} for the annotated attributes (subclasses and in-place edits included)."""
    names = {}
    for klass in reversed(hparams_class.__mro__):
        names.update(getattr(klass, "__annotations__", {}))
    return {attr: getattr(hparams_class, attr) for attr in names}


def print_hparams(hparams_class) -> None:
    for attr, value in hparams_to_dict(hparams_class).items():
        print(f"{attr}: {value}")
