"""Import-path compatibility with the reference's ``bnn.binarize``; the code lives in ``convert``."""
from .convert import *  # noqa: F401,F403
from .convert import (_KNOWN_SPECIAL_WORDS, _get_first_layer, _get_last_layer,  # noqa: F401
                      _regex_match)
