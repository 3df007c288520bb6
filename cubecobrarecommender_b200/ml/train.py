"""Drop-in for reference ``src/ml/train.py``.

    python -m cubecobrarecommender_b200.ml.train epochs batch_size name reg noise [seed]

Same positional arguments (reference train.py:28-38), same inputs (``data/maps/nameToId.json``,
``data/cube/*.json``, ``output/full_adj_mtx.npy``), same M-hat construction (train.py:69-71), loss and
optimiser (train.py:83-88).  The model is saved under ``ml_files/<name>/`` in this package's npz
checkpoint format (TensorFlow SavedModel is not available).  Under ``torchrun`` the cubes
of every batch and the regulariser rows are sharded over the ranks and gradients are all-reduced.
"""
from __future__ import annotations

import os
import random
import sys
import time

import numpy as np
import torch


def reset_random_seeds(seed):
    os.environ['PYTHONHASHSEED'] = str(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def fit(model, generator, epochs, reg, *, group=None, log=print, steps_per_epoch=None, max_cube_size=None,
        initial_epoch=0, checkpoint_dir=None, save_every=None, overflow_check_every=50, metrics=True):
    """``autoencoder.fit(generator, epochs=epochs)`` (reference train.py:99-102).  Returns the history
    ``[{"loss", "output_1_loss", "output_2_loss", "output_1_accuracy", "output_2_accuracy"}]`` per epoch (means over
    the epoch's steps, like Keras' progress line).  ``metrics`` mirrors ``compile(..., metrics=['accuracy'])`` (reference
    train.py:87): Keras resolves the name to binary accuracy on the sigmoid output and categorical accuracy on the
    softmax output; both are counted inside the loss kernels.

    Beyond the reference: ``initial_epoch`` (Keras' own argument name) continues a run -- the epoch
    shuffle is a function of (seed, epoch), Adam's step counter lives in the model -- and every ``save_every`` epochs the
    model, its Adam slots and the number of completed epochs are written to ``checkpoint_dir`` (rank 0 writes; in p2p
    data-parallel mode the sliced Adam slots are gathered first)."""
    import torch.distributed as dist
    from .engine import DAEEngine
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    gb = generator.batch_size
    if gb % world:
        raise ValueError(f"batch_size {gb} must be divisible by the number of ranks {world}")
    lb = gb // world
    eng = DAEEngine(model, generator.mhat_device(), batch=lb, reg_rows=lb, reg=reg,
                    max_cube_size=max_cube_size or max(generator.csr.max_size, 1), global_batch=gb,
                    global_reg_rows=gb, group=group, metrics=metrics)
    history = []
    nsteps = steps_per_epoch or len(generator)
    if initial_epoch:
        generator.reset_indices(initial_epoch)
    for epoch in range(initial_epoch, epochs):
        log(f"Epoch {epoch + 1}/{epochs}")
        t0 = time.time()
        acc = torch.zeros(3, dtype=torch.float64, device=model.device)
        macc = torch.zeros(2, dtype=torch.float64, device=model.device)
        for b in range(nsteps):
            ids_t = generator.batch_ids(b, rank, world)           # a slice of the epoch's device-side permutation
            eng.sample_batch(generator.indptr, generator.indices_dev, ids_t, generator.alias_prob, generator.alias_idx,
                             generator.noise, generator.noise_std, seed=generator.seed + 7919 * rank)
            acc += eng.train_step()
            if metrics:
                macc += eng.metrics2
            if overflow_check_every and (b + 1) % overflow_check_every == 0:
                eng.check_overflow()                                # (one 4-byte read: the only host sync inside an epoch)
        eng.check_overflow()
        mean = (acc / max(nsteps, 1)).cpu().numpy()
        history.append({"loss": float(mean[2]), "output_1_loss": float(mean[0]), "output_2_loss": float(mean[1])})
        line = (f"{nsteps}/{nsteps} - {time.time() - t0:.0f}s - loss: {mean[2]:.4f} - output_1_loss: {mean[0]:.4f} "
                f"- output_2_loss: {mean[1]:.4f}")
        if metrics:
            mm = (macc / max(nsteps, 1)).cpu().numpy()
            history[-1].update({"output_1_accuracy": float(mm[0]), "output_2_accuracy": float(mm[1])})
            line += f" - output_1_accuracy: {mm[0]:.4f} - output_2_accuracy: {mm[1]:.4f}"
        log(line)
        generator.on_epoch_end()
        if checkpoint_dir and save_every and (epoch + 1) % save_every == 0 and epoch + 1 < epochs:
            eng.gather_adam_state()
            if rank == 0:
                model.save(checkpoint_dir, epoch=epoch + 1)
    eng.gather_adam_state()      # p2p data parallel: m / v live sliced over the ranks until a checkpoint needs them
    return history


def main(argv=None):
    """``train.py epochs batch_size name reg noise [seed]`` (reference train.py:28-38), plus two optional flags the
    reference does not have: ``--save-every N`` writes a checkpoint to ``ml_files/<name>`` every N epochs, ``--resume``
    continues from the checkpoint found there (weights, Adam slots, step counter, completed epochs)."""
    from ..non_ml import utils
    from .generator import DataGenerator
    from .model import CC_Recommender
    args = list(sys.argv[1:] if argv is None else argv)
    resume = "--resume" in args
    if resume:
        args.remove("--resume")
    save_every = None
    if "--save-every" in args:
        i = args.index("--save-every")
        save_every = int(args[i + 1])
        del args[i:i + 2]
    epochs, batch_size, name = int(args[0]), int(args[1]), args[2]
    reg, noise = float(args[3]), float(args[4])
    seed = 0
    if len(args) == 6:
        seed = int(args[5])
        reset_random_seeds(seed)
    import torch.distributed as dist
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    map_file = '././data/maps/nameToId.json'
    folder = "././data/cube/"
    print('Loading Cube Data . . .\n')
    num_cards, name_lookup, card_to_int, int_to_card = utils.get_card_maps(map_file)
    cubes = utils.build_cubes_csr(folder, num_cards, name_lookup, card_to_int)
    print('Loading Adjacency Matrix . . .\n')
    adj_mtx = np.load('././output/full_adj_mtx.npy')
    print('Creating Graph for Regularization . . . \n')
    y_mtx = adj_mtx.copy()
    np.fill_diagonal(y_mtx, 1)
    y_mtx = (y_mtx / y_mtx.sum(1)[:, None])                                   # train.py:69-71
    print('Setting Up Data for Training . . .\n')
    print('Setting Up Model . . . \n')
    dest = f'././ml_files/{name}'
    precision = os.environ.get("CC_PRECISION", "tf32")
    initial_epoch = 0
    if resume and os.path.exists(os.path.join(dest, "cc_recommender.npz")):
        autoencoder = CC_Recommender.load(dest, device="cuda", precision=precision)
        initial_epoch = autoencoder.completed_epochs
        print(f'Resuming from {dest}: {initial_epoch} epochs done, Adam step {int(autoencoder.store.step.item())}\n')
    else:
        autoencoder = CC_Recommender(num_cards, device="cuda", seed=seed, precision=precision)
    generator = DataGenerator(y_mtx, cubes, batch_size=batch_size, noise=noise, seed=seed)
    fit(autoencoder, generator, epochs, reg, initial_epoch=initial_epoch, checkpoint_dir=dest, save_every=save_every)
    if not dist.is_initialized() or dist.get_rank() == 0:
        autoencoder.save(dest, epoch=epochs)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
