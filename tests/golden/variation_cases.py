"""Case tables shared by make_golden_variation_rng.py (reference side) and
tests/test_oracle_variation.py (oracle side)."""
PATH_CASES = [(0, 150, 207, None), (1, 40, 1, None), (2, 199, 305, None), (3, 25, 10, 60),
              (4, 80, 207, 5), (5, 1, 0, None)]          # (seed, path_len, start, len of prev_path or None)
ORG_CASES = [(10, 0.3, 0.2), (11, 1.0, 0.0), (12, 0.0, 1.0), (13, 1.0, 1.0), (14, 1.0, 1.0),
             (15, 0.0, 0.0)]                              # (seed, mutation_prob, muterpolate_prob)
XOVER_SEEDS = [20, 21, 22, 23, 24, 25]                    # case 3 raises min_dna_len to 170 (padMinDNA)
EVO_CASES = [(30, 12, 3), (31, 8, 4)]                     # (seed, gen_size, generations)
