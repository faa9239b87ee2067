"""B200-native drop-in for the `matrix_factorization` package's KernelMF / BaselineModel path."""
