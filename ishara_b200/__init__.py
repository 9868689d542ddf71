"""ishara_b200 — B200-native (sm_100a) implementation of the Ishara landmark-encoder hot path."""
