from . import model_with_file, model_wrapper  # noqa: F401
