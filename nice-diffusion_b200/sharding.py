"""Batch sharding for multi-GPU sampling: one process per GPU, the images of every batch are split
over the ranks by GLOBAL sample index, no collective on the data path. Everything that depends on randomness is keyed by
that global index (Philox noise in the step kernel, labels / x_T drawn from a generator every rank seeds identically and
then slices), so the images do not depend on the number of GPUs."""
import os


def world_info():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def shard_range(total, rank, world):
    """Contiguous, balanced split of range(total): the first (total % world) ranks get one extra item."""
    assert 0 <= rank < world and total >= 0
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def global_indices(batch_index, batch_size, rank, world):
    """Global sample ids of this rank's slice of batch `batch_index` (range object)."""
    lo, hi = shard_range(batch_size, rank, world)
    return range(batch_index * batch_size + lo, batch_index * batch_size + hi)


def output_names(labels_per_image):
    """File names the reference gives the images in order: '{label}_sample{k}.jpg' with k counting per label, or
    'sample{k}.jpg' for unconditional models (sample.py:162-179). Computed from the full label list so every rank
    names its own images without talking to the others."""
    names, counts, n = [], {}, 0
    for lab in labels_per_image:
        if lab is None:
            names.append('sample{}.jpg'.format(n))
            n += 1
        else:
            k = counts.get(lab, 0)
            names.append('{}_sample{}.jpg'.format(lab, k))
            counts[lab] = k + 1
    return names
