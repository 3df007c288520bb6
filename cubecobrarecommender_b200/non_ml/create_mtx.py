"""Drop-in for reference ``src/non_ml/create_mtx.py``: same inputs, prints and output files
(``output/full_adj_mtx.npy`` float64 (C, C), ``output/int_to_card.json``).  Run from the data root:

    python -m cubecobrarecommender_b200.non_ml.create_mtx
"""
import json
import os
import os.path

import numpy as np

from . import utils


def main(root='.'):
    map_file = os.path.join(root, 'data/maps/nameToId.json')
    folder = os.path.join(root, 'data/cube/')
    print('getting data')
    num_cards, name_lookup, card_to_int, int_to_card = utils.get_card_maps(map_file)
    cubes = utils.build_cubes_csr(folder, num_cards, name_lookup, card_to_int)
    print('creating matrix')
    adj_mtx = utils.create_adjacency_matrix(cubes)
    dest = os.path.join(root, 'output')
    if not os.path.isdir(dest):
        os.makedirs(dest)
    with open(os.path.join(dest, 'full_adj_mtx.npy'), 'wb') as out_mtx:
        np.save(out_mtx, adj_mtx)
    with open(os.path.join(dest, 'int_to_card.json'), 'w') as out_lookup:
        json.dump(int_to_card, out_lookup)
    return adj_mtx


if __name__ == "__main__":
    main()
