class Error(Exception):
    pass


class NameNotFound(Error):
    pass


class ResetNeeded(Error):
    pass
