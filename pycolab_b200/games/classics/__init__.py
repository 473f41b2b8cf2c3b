"""Set-up twins of the reference's `pycolab/examples/classics/` games."""
