import tempfile


def string_to_file(string: str) -> str:
    """Write ``string`` to a fresh temporary file and return its path."""
    handle, path = tempfile.mkstemp()
    with open(handle, "w") as f:
        f.write(string)
    return path
