"""Search-path helper for model / world files."""
import os
from typing import List

_SEARCH_PATH: List[str] = []


def get_search_paths() -> List[str]:
    return list(_SEARCH_PATH)


def add_path(path: str) -> None:
    if os.path.isdir(path) and path not in _SEARCH_PATH:
        _SEARCH_PATH.append(path)


def add_path_from_env_var(env_variable: str) -> None:
    for p in os.environ.get(env_variable, "").split(":"):
        if p:
            add_path(p)


def find_resource(file_name: str) -> str:
    if os.path.isabs(file_name) and os.path.isfile(file_name):
        return file_name
    for base in [os.getcwd()] + _SEARCH_PATH:
        candidate = os.path.join(base, file_name)
        if os.path.isfile(candidate):
            return os.path.abspath(candidate)
    raise FileNotFoundError(f"Failed to find resource '{file_name}'")
