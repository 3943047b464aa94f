"""Stand-in for ``termcolor`` (utils/image_util.py:4): no colours."""


def colored(text, *args, **kwargs):
    return text


def cprint(text, *args, **kwargs):
    print(text)
