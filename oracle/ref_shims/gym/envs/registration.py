registry = {}


def register(id, entry_point, **_ignored):  # noqa: A002 - mirrors gym's keyword
    registry[id] = entry_point
