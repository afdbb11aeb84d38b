"""Stand-in for matplotlib (absent from this image): the reference's plots are observability only (SURVEY.md section
2); every call is accepted and draws nothing."""


def use(*args, **kwargs):
    pass
