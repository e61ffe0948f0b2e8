"""Stand-in for the `munch` package (not installable offline) so that the unmodified reference's
rectorch.configuration / rectorch.data can be imported by oracle/make_golden_data.py.  Only what
configuration.py:26-46 uses: DefaultMunch(default, mapping) = dict with attribute access that returns
`default` for a missing key."""


class Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class DefaultMunch(Munch):
    def __init__(self, *args, **kwargs):
        default = args[0] if args else None
        dict.__setattr__(self, "__default__", default)
        dict.__init__(self, *args[1:], **kwargs)

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return self.get(k, dict.__getattribute__(self, "__default__"))

    def __setattr__(self, k, v):
        self[k] = v
