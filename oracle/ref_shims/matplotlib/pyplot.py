class Axes(object):
    def __init__(self, *a, **k):
        pass


class NullFormatter(object):
    pass


class Polygon(object):
    def __init__(self, *a, **k):
        pass


def _noop(*a, **k):
    return None


figure = gcf = show = pause = close = plot = savefig = xlabel = ylabel = legend = \
    title = imshow = axis = subplots = _noop
