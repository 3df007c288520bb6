"""Drop-in for reference ``src/scripts/ml_recommend.py``.

    python -m cubecobrarecommender_b200.scripts.ml_recommend cube_id [N=100] [root]

Two arguments or fewer: prints the additions one per line, a blank line, then the N lowest-scored
in-cube cards with their scores (reference ml_recommend.py:98-116).  With a third argument the
reference builds the JSON dict but never prints it (:89-116); here it is printed as JSON.
"""
import json
import sys

import numpy as np


def main(argv=None, model_dir='ml_files/neg', id_map='ml_files/recommender_id_map.json'):
    from ..ml.inference import MLRecommender
    from ..ml.model import load_model
    from .common import cube_indices, fetch_cube_list, load_int_to_card
    args = sys.argv[1:] if argv is None else argv
    cube_name = args[0]
    non_json, root, amount = True, "https://cubecobra.com", 100
    if len(args) > 1:
        amount = int(args[1])
        if len(args) > 2:
            root, non_json = args[2], False
    print('Getting Cube List . . . \n')
    card_names = fetch_cube_list(cube_name, root)
    print('Loading Card Name Lookup . . . \n')
    int_to_card, card_to_int = load_int_to_card(id_map)
    print('Creating Cube Vector . . . \n')
    idxs = cube_indices(card_names, card_to_int)
    print('Loading Model . . . \n')
    rec = MLRecommender(load_model(model_dir))
    print('Generating Recommendations . . . \n')
    output = rec.recommend_one(idxs, amount, int_to_card)
    if non_json:
        for card in output['additions']:
            print(card)
        cards = list(output['cuts'].keys())
        vals = list(output['cuts'].values())
        rank_cuts = np.array(vals).argsort(kind='stable')
        print('\n')
        for i in rank_cuts[:amount]:
            print(cards[i], vals[i])
    else:
        print(json.dumps(output))
    return output


if __name__ == "__main__":
    main()
