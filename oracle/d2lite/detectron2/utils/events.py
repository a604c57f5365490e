from collections import defaultdict

_STACK = []


def get_event_storage():
    assert len(_STACK), "get_event_storage() has to be called inside a 'with EventStorage(...)' context!"
    return _STACK[-1]


class EventStorage:
    def __init__(self, start_iter=0):
        self.iter = start_iter
        self._history = defaultdict(list)
        self._latest = {}

    def put_scalar(self, name, value, smoothing_hint=True):
        value = float(value)
        self._history[name].append((value, self.iter))
        self._latest[name] = value

    def put_scalars(self, *, smoothing_hint=True, **kwargs):
        for k, v in kwargs.items():
            self.put_scalar(k, v, smoothing_hint)

    def put_image(self, name, img):
        pass

    def latest(self):
        return dict(self._latest)

    def history(self, name):
        return self._history[name]

    def step(self):
        self.iter += 1

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *a):
        assert _STACK[-1] is self
        _STACK.pop()
