"""Drop-in mirror of reference ``src/non_ml/utils.py`` (same names, arguments and return types).

``create_adjacency_matrix`` runs on the GPU through the C ABI
(``cc_create_adjacency_matrix_host``): bit-packed popcount counts + row normalise, bit-exact
with the reference's float64 result.  ``build_cubes_csr`` is the one-pass sparse sibling of
``build_cubes``: the reference parses every JSON twice and materialises a
dense float64 (K, C) matrix, which is what keeps it from scaling.
"""
from __future__ import annotations

import json
import os

import numpy as np

from ..sparse import CubeCSR


def exclude(card_file=None):
    """Reference ``utils.py:6-25``.  (The reference's ``cd.get['name_lower']`` at :24 is a latent
    TypeError that no caller reaches; here the lookup is the call it was meant to be.)"""
    if card_file is None:
        return []
    bad_names = ['plains', 'island', 'swamp', 'mountain', 'forest', '1996 world champion']
    card_dict = json.load(open(card_file, 'rb'))
    for cd in card_dict.values():
        if cd.get('isToken'):
            bad_names.append(cd.get('name_lower'))
    return bad_names


def get_card_maps(map_file, exclude_file=None):
    """Reference ``utils.py:27-47``: card ids follow the JSON's insertion order."""
    exclusions = exclude(exclude_file)
    names = json.load(open(map_file, 'rb'))
    name_lookup, card_to_int = dict(), dict()
    num_cards = 0
    for name, ids in names.items():
        if name in exclusions:
            continue
        card_to_int[name] = num_cards
        for idx in ids:
            name_lookup[idx] = name
        num_cards += 1
    int_to_card = {v: k for k, v in card_to_int.items()}
    return num_cards, name_lookup, card_to_int, int_to_card


def get_num_cubes(cube_folder):
    """Reference ``utils.py:49-55``."""
    num_cubes = 0
    for f in os.listdir(cube_folder):
        contents = json.load(open(os.path.join(cube_folder, f), 'rb'))
        num_cubes += len(contents)
    return num_cubes


def _cube_card_ids(cube, name_lookup, card_to_int):
    card_ids = []
    for card in cube['cards']:
        card_name = name_lookup.get(card['cardID'])
        if card_name is not None:
            card_id = card_to_int.get(card_name)
            if card_id is not None:
                card_ids.append(card_id)
    return card_ids


def build_cubes(cube_folder, num_cubes, num_cards, name_lookup, card_to_int):
    """Reference ``utils.py:57-73``: dense float64 (K, C), rows in ``os.listdir`` order."""
    cubes = np.zeros((num_cubes, num_cards))
    counter = 0
    for f in os.listdir(cube_folder):
        contents = json.load(open(os.path.join(cube_folder, f), 'rb'))
        for cube in contents:
            cubes[counter, _cube_card_ids(cube, name_lookup, card_to_int)] = 1
            counter += 1
    return cubes


def build_cubes_csr(cube_folder, num_cards, name_lookup, card_to_int) -> CubeCSR:
    """Same cubes, same order, one JSON pass, CSR output (duplicates collapse like ``cubes[i, ids] = 1``)."""
    lists = []
    for f in os.listdir(cube_folder):
        contents = json.load(open(os.path.join(cube_folder, f), 'rb'))
        for cube in contents:
            lists.append(_cube_card_ids(cube, name_lookup, card_to_int))
    return CubeCSR.from_lists(lists, num_cards)


def create_adjacency_matrix(cubes, verbose=True, force_diag=None):
    """Reference ``utils.py:75-92``.  ``cubes``: the dense 0/1 (K, C) array ``build_cubes`` returns, or a
    ``CubeCSR``.  Returns the float64 (C, C) matrix M[i, j] = P(j in cube | i in cube)."""
    from .. import graph
    csr = cubes if isinstance(cubes, CubeCSR) else CubeCSR.from_dense(np.asarray(cubes))
    num_cards = csr.num_cards
    if verbose:
        for i in range(0, num_cards, 100):
            print(i + 1, "/", num_cards)       # the reference's progress lines (utils.py:79-81)
    return graph.create_adjacency_matrix_host(csr, force_diag=force_diag)
