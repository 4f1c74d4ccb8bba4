"""gymnasium.utils.seeding.np_random restated: PCG64 seeded through numpy SeedSequence."""
import numpy as np

from .. import error

RandomNumberGenerator = np.random.Generator
RNG = np.random.Generator


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
        raise error.Error(f"Seed must be a python integer, actual type: {type(seed)}" if not isinstance(seed, int)
                          else f"Seed must be greater or equal to zero, actual value: {seed}")
    seed_seq = np.random.SeedSequence(None if seed is None else int(seed))
    np_seed = seed_seq.entropy
    rng = np.random.Generator(np.random.PCG64(seed_seq))
    return rng, np_seed
