"""Text front-end of the synthesis drivers: character ids and batch padding.

Mirrors generate_test_utterances.py:15-25 (text2id) and :56-73 (vocabulary table, zero padding
to the batch maximum).  Padding matters: the reference does not mask pad ids in TextEnc
(SURVEY.md F5), so K/V of a sentence depend on the padded length N of its batch.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np

DEFAULT_VOCABULARY = "PE abcdefghijklmnopqrstuvwxyz-,.?'\""     # reference config.json:12


def char_table(vocabulary: str = DEFAULT_VOCABULARY) -> dict:
    table = {ch: i for i, ch in enumerate(vocabulary)}
    table['"'] = len(vocabulary) - 2          # '"' shares the id of "'"
    return table


def text2id(text: str, vocabulary: str = DEFAULT_VOCABULARY, table: dict | None = None) -> np.ndarray:
    """Lower-case, append the end marker 'E', drop unknown characters -> int64 (1, len)."""
    table = table or char_table(vocabulary)
    ids = [table[ch] for ch in text.lower() + "E" if ch in vocabulary]
    return np.asarray(ids, dtype=np.int64).reshape(1, -1)


def pad_batch(rows: Sequence[np.ndarray], n: int | None = None) -> np.ndarray:
    """Right-pad id rows with 0 ('P') to length n (default: the longest row) -> int64 (U, n)."""
    longest = max(int(np.asarray(r).size) for r in rows)
    n = longest if n is None else n
    if n < longest:
        raise ValueError(f"pad length {n} shorter than the longest text ({longest})")
    out = np.zeros((len(rows), n), dtype=np.int64)
    for u, r in enumerate(rows):
        r = np.asarray(r).reshape(-1)
        out[u, : r.size] = r
    return out


def vocab_len(vocabulary: str = DEFAULT_VOCABULARY) -> int:
    """melSyn(vocab_len=...) as the drivers compute it (generate_test_utterances.py:75)."""
    return len(vocabulary) - 1


def encode_lines(lines: Iterable[str], vocabulary: str = DEFAULT_VOCABULARY) -> List[np.ndarray]:
    table = char_table(vocabulary)
    return [text2id(s.strip(), vocabulary, table) for s in lines]
