"""
Sequence_Magic.py — drop-in for the two helpers of the reference's Valkyries/Sequence_Magic.py that
sit on the hot path: match_maker (:123-136) and rcomp (:52-120).

match_maker(query, unknown) = Levenshtein.distance(query, unknown) - (len(unknown) - len(query)),
computed on the GPU by the bit-vector edit-distance kernel (scar_match_maker_batch); the batched form
match_maker_many is what to call from new code.  No CPU fallback.
"""
from __future__ import annotations

from .engine import match_maker_batch

_COMP = str.maketrans("ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn")


def rcomp(seq: str) -> str:
    """Reverse complement with the ambiguous-DNA table; unmapped characters pass through."""
    return seq[::-1].translate(_COMP)


def match_maker(query: str, unknown: str) -> int:
    return int(match_maker_batch([query], [unknown])[0])


def match_maker_many(queries, unknowns, device: int = 0):
    return match_maker_batch(queries, unknowns, device)
