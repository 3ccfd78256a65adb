"""``Token``: the vocabulary item type the trie is built from.

Behavioural mirror of the reference's ``genlm/backend/tokenization/token.py:9-90`` (a ``bytes`` subclass
carrying a token id; equality, hashing and ordering between Tokens go by id because several tokens may
share one byte string).  The tokenizer-decoding helpers of the reference (``decode_vocab``) are input
producers outside this hot path and are not rebuilt here.
"""


class Token(bytes):
    """A vocabulary token: its byte string plus the integer id that identifies it."""

    def __new__(cls, token_id, byte_string):
        if not isinstance(token_id, int):
            raise TypeError(f"token_id must be an int, got {type(token_id)}")
        if not isinstance(byte_string, bytes):
            raise TypeError(f"byte_string must be bytes, got {type(byte_string)}")
        self = super().__new__(cls, byte_string)
        self.token_id = token_id
        return self

    @property
    def byte_string(self):
        return bytes(self)

    def __repr__(self):
        return f"Token(token_id={self.token_id}, byte_string={bytes(self)!r})"

    # Token vs Token compares ids; Token vs anything else defers to the other operand
    def _cmp(self, other, op):
        if not isinstance(other, Token):
            return NotImplemented
        return op(self.token_id, other.token_id)

    def __eq__(self, other):
        return self._cmp(other, int.__eq__)

    def __ne__(self, other):
        return self._cmp(other, int.__ne__)

    def __lt__(self, other):
        return self._cmp(other, int.__lt__)

    def __le__(self, other):
        return self._cmp(other, int.__le__)

    def __gt__(self, other):
        return self._cmp(other, int.__gt__)

    def __ge__(self, other):
        return self._cmp(other, int.__ge__)

    def __hash__(self):
        return hash(self.token_id)

    @staticmethod
    def as_bytes(x):
        return x.byte_string if isinstance(x, Token) else x

    @staticmethod
    def is_plain_bytes(x):
        return isinstance(x, bytes) and not isinstance(x, Token)

    def __reduce__(self):
        return (Token, (self.token_id, bytes(self)))
