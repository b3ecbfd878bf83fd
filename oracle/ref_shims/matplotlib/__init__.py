"""Stand-in for matplotlib (absent from this image): only pyplot.ion / pyplot.imsave are used by
the reference (common/generator.py:17,466-469)."""
