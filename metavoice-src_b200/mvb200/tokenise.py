"""Host-side BPE wrapper with the interface of ``fam/quantiser/text/tokenise.py:4-32`` (tiktoken core).
Text tokenisation is microseconds of host work and stays above the kernel boundary."""
from __future__ import annotations

import tiktoken


class TrainedBPETokeniser:
    def __init__(self, name, pat_str, mergeable_ranks, special_tokens, offset=None) -> None:
        self.tokenizer = tiktoken.Encoding(name=name, pat_str=pat_str, mergeable_ranks=mergeable_ranks,
                                           special_tokens=special_tokens)
        self.offset = offset

    def _shift(self, ids, sign):
        return ids if self.offset is None else [i + sign * self.offset for i in ids]

    def encode(self, text: str) -> list[int]:
        # an end-of-text token is always appended (reference tokenise.py:15-16)
        return self._shift(self.tokenizer.encode(text) + [self.tokenizer.eot_token], +1)

    def decode(self, tokens: list[int]) -> str:
        return self.tokenizer.decode(self._shift(list(tokens), -1))

    @property
    def eot_token(self) -> int:
        return self.tokenizer.eot_token + (self.offset or 0)
