"""Weight-file plumbing: a dependency-free reader for Keras HDF5 weight files."""
