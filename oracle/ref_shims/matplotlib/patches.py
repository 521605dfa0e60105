class Rectangle(object):
    def __init__(self, *a, **k):
        pass


class Polygon(object):
    def __init__(self, *a, **k):
        pass
