"""Stand-in: names imported by /root/reference/src/gcm/gcm.py:3 (unused on the hot path)."""


class Data:  # pragma: no cover
    pass


class Batch:  # pragma: no cover
    pass
