"""Mirror of spotlight/helpers.py:1-12."""


def _repr_model(model):
    inner = '[uninitialised]' if model._net is None else repr(model._net)
    return '<%s: %s>' % (type(model).__name__, inner)
