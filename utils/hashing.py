"""``utils.hashing`` -- the module the reference imports its retrieval evaluation from
(``experiments/test_hashing.py:15``, ``experiments/train_helper.py:18``; also ``get_hamm_dist`` in
``trainers/orthohash.py:16`` / ``trainers/dpn.py:4`` and the loss helpers ``get_sim`` / ``log_trick`` in
``models/loss/{dpsh,hashnet,adsh}.py``) and does not ship
(``README.md:11`` points at another repository).  This file is the binding a maintainer drops into the
reference's ``utils/`` directory: it forwards to the B200-native implementation.

``utils`` is a namespace package here on purpose (no ``__init__.py``), exactly like the reference's own
``utils/`` directory, so ``utils.metrics`` of the reference keeps resolving when both trees are on
``sys.path``.
"""
from concepthash_b200.codes_io import PackedCodes, evaluate_dumps, load_packed, save_packed  # noqa: F401
from concepthash_b200.hashing import (  # noqa: F401
    calculate_mAP,
    calculate_pr_curve,
    get_hamm_dist,
    get_sim,
    log_trick,
    map_at_r,
    pack_codes,
    retrieve_topk,
)

__all__ = ["calculate_mAP", "calculate_pr_curve", "get_hamm_dist", "get_sim", "log_trick", "map_at_r", "retrieve_topk",
           "PackedCodes", "evaluate_dumps", "load_packed", "save_packed", "pack_codes"]
