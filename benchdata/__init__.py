"""Synthetic workloads of the BASELINE.json configurations (SURVEY.md §8(d) generator).  Bench / test data only:
nothing here imports the product package, so the CPU reference arm of bench.py can use it without mapping the CUDA
library."""
