import sys

DEBUG, INFO, WARN, ERROR, DISABLED = 10, 20, 30, 40, 50
MIN_LEVEL = 30


def set_level(level):
    global MIN_LEVEL
    MIN_LEVEL = level


def _emit(level, tag, msg, args):
    if MIN_LEVEL <= level:
        print(f"{tag}: {msg % args if args else msg}", file=sys.stderr)


def debug(msg, *args):
    _emit(DEBUG, "DEBUG", msg, args)


def info(msg, *args):
    _emit(INFO, "INFO", msg, args)


def warn(msg, *args):
    _emit(WARN, "WARN", msg, args)


def error(msg, *args):
    _emit(ERROR, "ERROR", msg, args)
