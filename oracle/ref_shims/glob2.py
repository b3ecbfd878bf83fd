"""Stand-in for glob2.glob (absent from this image)."""
import glob as _glob


def glob(pattern):
    return _glob.glob(pattern, recursive=True)
