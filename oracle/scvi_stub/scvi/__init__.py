"""TEST INFRASTRUCTURE ONLY — stand-in for scvi-tools 0.20.0 (pinned by the reference's
pyproject.toml:25, source not available offline).  Restates, from the published
behaviour of that release, the five symbols the reference hot path imports
(reference: src/spVIPES/module/spVIPESmodule.py:7-9, src/spVIPES/nn/networks.py:5)
so that those files can be imported UNMODIFIED to generate golden vectors.

parity unpinned: no reference test or golden vector pins these behaviours.
Never imported by the product package (spvipes_b200/).
"""
from types import SimpleNamespace

REGISTRY_KEYS = SimpleNamespace(
    X_KEY="X", BATCH_KEY="batch", LABELS_KEY="labels", PROTEIN_EXP_KEY="proteins",
    CAT_COVS_KEY="extra_categorical_covs", CONT_COVS_KEY="extra_continuous_covs",
    INDICES_KEY="ind_x",
)
settings = SimpleNamespace(seed=0, batch_size=128, dl_pin_memory_gpu_training=False)
