"""rand 0.8 `StdRng` restated, so that "the same seeded RNG" gives the same (r, s) as the reference prover.

`create_random_proof_with_reduction` (forks/groth16/src/prover.rs:142-154) draws `r = Fr::rand(rng)` then `s = Fr::rand(rng)`;
Crescent passes `thread_rng()` (creds/src/lib.rs:281), its tests and benches `ark_std::test_rng()` / `StdRng::seed_from_u64`
(creds/src/rangeproof.rs:490-511, creds/benches/proof_benchmark.rs:85-96).  All of them are ChaCha12 block generators:

  * rand 0.8 `StdRng` = rand_chacha 0.3 `ChaCha12Rng`: the original (djb) ChaCha layout -- 4 constant words, 8 key words
    (the 32-byte seed, little-endian), a 64-bit block counter in words 12-13 starting at 0, a 64-bit stream id (0) in words
    14-15 -- with 12 rounds; the output is the keystream read as little-endian u32 words;
  * rand_core `BlockRng::next_u64`: two consecutive u32 words (low word first) from a 64-word buffer (4 blocks), with the
    straddling rule at the end of the buffer restated in `next_u64` below;
  * `SeedableRng::seed_from_u64`: the seed bytes come from a PCG32 stream (rand_core 0.6);
  * `ark_std::test_rng()`: `StdRng::from_seed([1,0,0,0, 23,0,0,0, 200,1,0,0, 210,30,0,0, 0 x 16])` (ark-std 0.4).

ASSUMPTION (third-party crates, not vendored in the reference tree): the constants above are restated from the published
sources of rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4 / ark-std 0.4.0.  The block function itself is pinned by the
RFC 8439 section 2.3.2 vector (20 rounds) in tests/test_host_logic.py; the 12-round count is the only difference.
"""
from __future__ import annotations

import struct
from typing import List

_MASK32 = 0xFFFFFFFF
_SIGMA = (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)  # "expand 32-byte k"


def _rotl(x: int, n: int) -> int:
    return ((x << n) | (x >> (32 - n))) & _MASK32


def _quarter(s: List[int], a: int, b: int, c: int, d: int) -> None:
    s[a] = (s[a] + s[b]) & _MASK32
    s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & _MASK32
    s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & _MASK32
    s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & _MASK32
    s[b] = _rotl(s[b] ^ s[c], 7)


def chacha_block(state: List[int], rounds: int) -> List[int]:
    """One ChaCha block: `state` is the 16-word input, the result the 16 output words (input added back)."""
    w = list(state)
    for _ in range(rounds // 2):
        _quarter(w, 0, 4, 8, 12)
        _quarter(w, 1, 5, 9, 13)
        _quarter(w, 2, 6, 10, 14)
        _quarter(w, 3, 7, 11, 15)
        _quarter(w, 0, 5, 10, 15)
        _quarter(w, 1, 6, 11, 12)
        _quarter(w, 2, 7, 8, 13)
        _quarter(w, 3, 4, 9, 14)
    return [(x + y) & _MASK32 for x, y in zip(w, state)]


class ChaChaRng:
    """rand_chacha `ChaChaXRng` behind rand_core's `BlockRng` (64-word buffer = 4 blocks)."""
    ROUNDS = 12
    BUF_WORDS = 64

    def __init__(self, seed: bytes, stream: int = 0):
        if len(seed) != 32:
            raise ValueError("seed must be 32 bytes")
        self.key = list(struct.unpack("<8I", seed))
        self.stream = stream
        self.counter = 0                 # block counter of the next block to generate
        self.results: List[int] = [0] * self.BUF_WORDS
        self.index = self.BUF_WORDS      # empty buffer

    @classmethod
    def from_seed(cls, seed) -> "ChaChaRng":
        return cls(bytes(seed))

    @classmethod
    def seed_from_u64(cls, state: int) -> "ChaChaRng":
        """rand_core 0.6 SeedableRng::seed_from_u64: 8 PCG32 outputs, little-endian."""
        mul, inc, m64 = 6364136223846793005, 11634580027462260723, (1 << 64) - 1
        out = bytearray()
        for _ in range(8):
            state = (state * mul + inc) & m64
            xorshifted = (((state >> 18) ^ state) >> 27) & _MASK32
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & _MASK32
            out += struct.pack("<I", x)
        return cls(bytes(out))

    def _generate(self) -> None:
        buf: List[int] = []
        for _ in range(self.BUF_WORDS // 16):
            st = list(_SIGMA) + self.key + [self.counter & _MASK32, (self.counter >> 32) & _MASK32,
                                            self.stream & _MASK32, (self.stream >> 32) & _MASK32]
            buf += chacha_block(st, self.ROUNDS)
            self.counter = (self.counter + 1) & ((1 << 64) - 1)
        self.results = buf

    def next_u32(self) -> int:
        if self.index >= self.BUF_WORDS:
            self._generate()
            self.index = 0
        v = self.results[self.index]
        self.index += 1
        return v

    def next_u64(self) -> int:
        n = self.BUF_WORDS
        i = self.index
        if i < n - 1:
            self.index += 2
            return self.results[i] | (self.results[i + 1] << 32)
        if i >= n:
            self._generate()
            self.index = 2
            return self.results[0] | (self.results[1] << 32)
        lo = self.results[n - 1]      # one word left: it is the low half, the first word of the next buffer the high half
        self._generate()
        self.index = 1
        return lo | (self.results[0] << 32)

    def getrandbits(self, k: int) -> int:
        """random.Random-style accessor used by groth16.sample_fr (k = 64: next_u64, k = 32: next_u32)."""
        if k == 64:
            return self.next_u64()
        if k == 32:
            return self.next_u32()
        raise ValueError("only 32- or 64-bit draws are defined for the Rust RNG mirror")


class StdRng(ChaChaRng):
    """rand 0.8 StdRng (ChaCha12)."""
    ROUNDS = 12


class ChaCha20Rng(ChaChaRng):
    ROUNDS = 20


def test_rng() -> StdRng:
    """ark_std::test_rng() (ark-std 0.4, std feature off / deterministic build)."""
    seed = [1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16
    return StdRng.from_seed(seed)
