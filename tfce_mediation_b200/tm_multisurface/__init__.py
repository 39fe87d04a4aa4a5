"""mmr-lr permutation path (reference: tm_multisurface/tm_mmr_rand_low_ram*.py), batched over
shuffles x surfaces on the GPU."""
