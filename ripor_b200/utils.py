"""Output-format helpers of the retrieval path (reference t5_pretrainer/utils/utils.py)."""
from __future__ import annotations

from typing import List


def get_dataset_name(path: str) -> str:
    """Dataset name from a query-collection or qrel path; same table as reference utils.py:13-35."""
    table = (("TREC_DL_2019", "TREC_DL_2019"), ("trec2020", "TREC_DL_2020"), ("TREC_DL_2020", "TREC_DL_2020"))
    for needle, name in table:
        if needle in path:
            return name
    if "msmarco" in path:
        return "MSMARCO_TRAIN" if "train_queries" in path else "MSMARCO"
    if "MSMarco-v2" in path:
        if "dev_1" in path:
            return "MSMARCO_v2_dev1"
        assert "dev_2" in path
        return "MSMARCO_v2_dev2"
    if "toy" in path:
        return "TOY"
    if "nq-320k" in path:
        return "NQ_320K"
    return "other_dataset"


def convert_ptsmtids_to_strsmtid(input_smtids, seq_length: int) -> List[List[str]]:
    """[B, nb, L+1] token ids -> "c1_c2_.._cL" strings, dropping the decoder start column (utils.py:46-59)."""
    assert input_smtids.dim() == 3, input_smtids.dim()
    assert input_smtids.size(2) == seq_length + 1, (input_smtids.size(1), seq_length)
    return [["_".join(str(x) for x in smtids[1:]) for smtids in beams] for beams in input_smtids.cpu().tolist()]
