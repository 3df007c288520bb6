"""Shared pieces of the CLI drop-ins: cube list fetch, name normalisation, cube vector."""
from __future__ import annotations

import json
import unicodedata
import urllib.request

import numpy as np

try:  # reference requirement; not installed in the build image
    from unidecode import unidecode as _unidecode
except ImportError:  # pragma: no cover
    _EXTRA = {"æ": "ae", "Æ": "AE", "œ": "oe", "Œ": "OE", "ß": "ss", "ø": "o", "Ø": "O", "đ": "d", "Đ": "D",
              "ł": "l", "Ł": "L", "þ": "th", "Þ": "Th", "’": "'", "‘": "'", "“": '"', "”": '"', "–": "-", "—": "--"}

    def _unidecode(s: str) -> str:
        s = "".join(_EXTRA.get(ch, ch) for ch in s)
        return "".join(ch for ch in unicodedata.normalize("NFKD", s) if not unicodedata.combining(ch)) \
            .encode("ascii", "ignore").decode("ascii")


def normalise(name: str) -> str:
    """``unidecode.unidecode(name.lower())`` (reference recommend.py:53)."""
    return _unidecode(name.lower())


def fetch_cube_list(cube_name: str, root: str = "https://cubecobra.com"):
    """Reference recommend.py:29-37: newline separated card names of a CubeCobra cube."""
    url = root + "/cube/api/cubelist/" + cube_name
    fp = urllib.request.urlopen(url)
    mystr = fp.read().decode("utf8")
    fp.close()
    return mystr.split("\n")


def load_int_to_card(path):
    """``{"<int>": name}`` JSON -> (int_to_card, card_to_int) (reference recommend.py:45-47)."""
    int_to_card = json.load(open(path, 'r'))
    int_to_card = {int(k): v for k, v in int_to_card.items()}
    card_to_int = {v: k for k, v in int_to_card.items()}
    return int_to_card, card_to_int


def cube_indices(card_names, card_to_int):
    """Reference recommend.py:49-56: unknown cards (custom cards) are skipped."""
    out = []
    for name in card_names:
        idx = card_to_int.get(normalise(name))
        if idx is not None:
            out.append(idx)
    return out


def cube_vector(indices, num_cards):
    cube = np.zeros(num_cards)
    cube[indices] = 1
    return cube
