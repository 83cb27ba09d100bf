"""Import alias: the product package lives in ``desta2.5-audio_b200/`` (not a legal Python identifier),
``import desta_b200`` resolves its submodules there."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "desta2.5-audio_b200")
__path__.insert(0, _impl)
exec(compile(open(_os.path.join(_impl, "__init__.py")).read(), _os.path.join(_impl, "__init__.py"), "exec"))
