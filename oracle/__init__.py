"""oracle -- TEST INFRASTRUCTURE ONLY (CPU checkers for the CUDA path). Never imported by the product package."""
