"""Stock-PyTorch stand-in for the slice of torch_xla the reference imports (see baseline/README.md)."""
