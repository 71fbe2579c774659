"""Test inputs shared by the CPU (oracle / simulator) and GPU parity tests.

`ROUNDTRIP_BATCH` are the strings of the reference's own `roundtrip_batch` (test/test/util.hpp:98-178) and
`generator_strings` follows `on_string_generators` (test/test/util.hpp:180-207): fib / thue-morse / run-rich words
0..n-1 and random digit strings of length 2^i.  Every case is returned as the escaped text + sentinel
(`test::compress_input`, test/test/util.hpp:507-539).
"""
import numpy as np

from tudocomp_b200 import synth

ROUNDTRIP_BATCH = [
    "abcdebcdeabc", "a", "", "aaaaaaaaa", "banana", "ananas", "abcdefgh#defgh_abcde", "abcdebcdeabcd", "foobar",
    "abcabcabcabc", "abc abc  abc", "abaaabbababb",
    "asdfasctjkcbweasbebvtiwetwcnbwbbqnqxernqzezwuqwezuetqcrnzxbneqebwcbqwicbqcbtnqweqxcbwuexcbzqwezcqbwecqbw"
    "dassdasdfzdfgfsdfsdgfducezctzqwebctuiqwiiqcbnzcebzqc",
    "ประเทศไทย中华Việt Nam",
    "Lorem ipsum dolor sit amet, sea ut etiam solet salutandi, sint complectitur et his, ad salutandi imperdi"
    "et gubergren per mei.",
    "Лорэм атоморюм ут хаж, эа граэки емпыдит ёудёкабет мэль, декам дежпютатионй про ты. Нэ ёужто жэмпэр"
    " жкрибэнтур векж, незл коррюмпит.",
    "報チ申猛あち涙境ワセ周兵いわ郵入せすをだ漏告されて話巡わッきや間紙あいきり諤止テヘエラ鳥提フ健2銀稿97傷エ映田ヒマ役請多"
    "暫械ゅにうて。関国ヘフヲオ場三をおか小都供セクヲ前俳著ゅ向深まも月10言スひす胆集ヌヱナ賀提63劇とやぽ生牟56詰ひめつそ総愛"
    "ス院攻せいまて報当アラノ日府ラのがし。",
    "Εαμ ανσιλλαε περισυλα συαφιθαθε εξ, δυο ιδ ρεβυμ σομμοδο. Φυγιθ ηομερω ιυς ατ, ει αυδιρε ινθελλεγαμ νες."
    " Ρεκυε ωμνιυμ μανδαμυς κυο εα. Αδμοδυμ σωνσεκυαθ υθ φιξ, εσθ ετ πρωβατυς συαφιθαθε ραθιονιβυς, ταντας αυ"
    "διαμ ινστρυσθιορ ει σεα.",
    "struct Foo { uint8_t bar }", "ABBCBCABA", "abcabca", "abbbbbbbbbbcbbbbbbbbbb",
]


def roundtrip_batch():
    for i, s in enumerate(ROUNDTRIP_BATCH):
        yield f"batch{i}", synth.escape_with_sentinel(s.encode("utf-8"))
    # escaping golden bytes of test/tudocomp_tests.cpp:528-533: "\0\x01\xff\xfe\0" -> ff fe 01 ff ff fe ff fe 00
    yield "escape_kat", synth.escape_with_sentinel(b"\x00\x01\xff\xfe\x00")
    yield "all_bytes", synth.escape_with_sentinel(bytes(range(256)))


def generator_strings(n=11):
    for i in range(n):
        yield f"fib{i}", synth.escape_with_sentinel(synth.fib_word(i))
    for i in range(n):
        yield f"thue_morse{i}", synth.escape_with_sentinel(synth.thue_morse_word(i))
    for i in range(n):
        yield f"run_rich{i}", synth.escape_with_sentinel(synth.run_rich_word(i))
    for i in range(2, n):
        for j in range(1, 3):
            yield f"random{i}_{j}", synth.escape_with_sentinel(synth.random_digits(1 << i, j))


def small_synthetic():
    """Seeded small versions of the BASELINE.json workloads + adversarial shapes (multi-tile, multi-round)."""
    yield "dna_20k", synth.dna(20000, 5)
    yield "markov_30k", synth.markov_text(30000, 7)
    yield "repetitive_40k", synth.repetitive(40000, 3, block=1000, p=0.01)
    yield "run_a_9000", synth.with_sentinel(np.full(9000, 97, np.uint8))
    rng = np.random.default_rng(1)
    yield "bytes_15k", synth.with_sentinel(rng.integers(1, 255, 15000, dtype=np.uint8))
    yield "fib20", synth.escape_with_sentinel(synth.fib_word(20))
    yield "period7_10k", synth.with_sentinel(np.tile(np.frombuffer(b"abcabdx", np.uint8), 1500))


def all_small_cases():
    yield from roundtrip_batch()
    yield from generator_strings()
    yield from small_synthetic()
