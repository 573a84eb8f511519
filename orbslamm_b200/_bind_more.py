"""ctypes declarations of the matcher / optimizer entry points (filled in as they are added)."""


def declare(L):
    pass
