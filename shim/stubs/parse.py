"""Minimal stand-in for the `parse` package (format-string inverse), enough for the reference's file-name bookkeeping
(train.py:117-118: parse('{prefix}_{ID}.{ext}', file).named['ID'])."""
import re


class Result:
    def __init__(self, named):
        self.named = named

    def __getitem__(self, key):
        return self.named[key]


def parse(fmt, string):
    pattern, pos = '', 0
    for m in re.finditer(r'\{(\w*)(?::[^}]*)?\}', fmt):
        pattern += re.escape(fmt[pos:m.start()])
        pattern += f'(?P<{m.group(1)}>.+?)' if m.group(1) else '(.+?)'
        pos = m.end()
    pattern += re.escape(fmt[pos:])
    m = re.fullmatch(pattern, string)
    return Result(m.groupdict()) if m else None
