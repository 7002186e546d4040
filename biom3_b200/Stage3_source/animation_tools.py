"""Mirror of the on-path part of /root/reference/Stage3_source/animation_tools.py (:7-11)."""
from __future__ import annotations


def convert_num_to_char(tokens: list, char_tokens) -> str:
    """ids -> symbols, concatenated (animation_tools.py:7-11). ``char_tokens``: 1-D array of ids."""
    ids = char_tokens.tolist() if hasattr(char_tokens, 'tolist') else list(char_tokens)
    return ''.join(tokens[i] for i in ids)
