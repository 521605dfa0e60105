class Button(object):
    def __init__(self, *a, **k):
        pass
