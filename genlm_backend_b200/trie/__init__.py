from .base import TokenCharacterTrie
from .parallel import ParallelTokenCharacterTrie
from .async_impl import AsyncTokenCharacterTrie

__all__ = ["TokenCharacterTrie", "ParallelTokenCharacterTrie", "AsyncTokenCharacterTrie"]
