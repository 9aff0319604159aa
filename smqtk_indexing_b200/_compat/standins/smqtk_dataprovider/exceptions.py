"""Stand-in for ``smqtk_dataprovider.exceptions``."""


class ReadOnlyError(Exception):
    """Raised when a mutation is requested of a read-only structure."""


class InvalidUriError(Exception):
    def __init__(self, uri_value: str, reason: str):
        super().__init__(uri_value, reason)
        self.uri = uri_value
        self.reason = reason
