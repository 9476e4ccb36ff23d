"""Known answers transcribed from the reference's own test modules (paths relative to /root/reference).
Each case: (name, input values, [(op, needle, expected mask)]). `None` in a mask = null. Used twice: to pin
the CPU oracle (tests/test_oracle_golden.py, no GPU) and to check the CUDA path (tests/test_gpu_golden.py).
ops: "=", "!=", "<", "<=", ">", ">=" ; "like"/"not like" take the SQL pattern with its % signs and need
fingerprints (hint SubstringSearch) unless `fp` is False in the case name dict.
"""

LONG_A = "prefix_" + "a" * 253
LONG_B = "prefix_" + "b" * 253

# src/core/src/liquid_array/byte_view_array/tests.rs
BYTE_VIEW_CASES = [
    ("shared_prefix_functionality (tests.rs:190-226)", ["hello_world", "hello_rust", "hello_test", "hello_code"], [
        ("=", "hello_rust", [False, True, False, False]),
        ("=", "goodbye_world", [False, False, False, False]),
        ("=", "hello_", [False, False, False, False]),
    ]),
    ("shared_prefix_with_short_strings (tests.rs:228-264)", ["abc", "abcde", "abcdef", "abcdefg"], [
        ("=", "abc", [True, False, False, False]),
        ("=", "abcde", [False, True, False, False]),
        (">", "ab", [True, True, True, True]),
        ("<", "abcd", [True, False, False, False]),
    ]),
    ("shared_prefix_contains_complete_strings (tests.rs:266-315)", ["data", "database", "data_entry", "data_", "datatype"], [
        ("=", "data", [True, False, False, False, False]),
        (">", "dat", [True, True, True, True, True]),
        ("<", "datab", [True, False, True, True, False]),
        (">", "da", [True, True, True, True, True]),
        (">=", "data", [True, True, True, True, True]),
        (">", "data", [False, True, True, True, True]),
    ]),
    ("prefix_optimization_fast_path (tests.rs:478-510)", ["apple123", "banana456", "cherry789", "apple999", "zebra000"], [
        ("<", "car", [True, True, False, True, False]),
        (">", "dog", [False, False, False, False, True]),
        (">=", "apple", [True, True, True, True, True]),
    ]),
    ("prefix_optimization_decompression_path (tests.rs:512-542)",
     ["prefix_aaa", "prefix_bbb", "prefix_ccc", "prefix_abc", "different"], [
        ("<", "prefix_b", [True, False, False, True, True]),
        ("<=", "prefix_bbb", [True, True, False, True, True]),
        (">", "prefix_abc", [False, True, True, False, False]),
    ]),
    ("prefix_optimization_edge_cases_and_nulls (tests.rs:544-616)", ["", None, "a", "abcdef", "abcdefghij", "abcdeg"], [
        ("<", "", [False, None, False, False, False, False]),
        (">", "abcdef", [False, None, False, False, True, True]),
        ("<=", "b", [True, None, True, True, True, True]),
        (">=", "abcdeg", [False, None, False, False, False, True]),
    ]),
    ("prefix_empty_suffix (tests.rs:618-631)", ["x", "x1"], [
        ("<=", "x", [True, False]),
        (">", "x", [False, True]),
    ]),
    ("utf8_and_binary (tests.rs:633-685)", ["café", "naïve", "résumé", "hello", "世界"], [
        ("<", "naïve", [True, False, False, True, False]),
        (">", "café", [False, True, True, True, True]),
        ("<=", "世界", [True, True, True, True, True]),
        (">=", "résumé", [False, False, True, False, True]),
        ("<=", "résumé", [True, True, True, True, False]),
    ]),
    ("compare_equals_on_disk inputs, in-memory semantics (tests.rs:689-751)",
     ["apple_orange", None, "apple_orange_long_string", "apple_b", "apple_oo_long_string", "apple_b", "apple"], [
        ("=", "apple", [False, None, False, False, False, False, True]),
        ("=", "", [False, None, False, False, False, False, False]),
        ("=", "apple_b", [False, None, False, True, False, True, False]),
        ("=", "apple_oo_long_string", [False, None, False, False, True, False, False]),
    ]),
    ("long_string_len_byte_255 (tests.rs:753-774)", [LONG_A, LONG_B, "z"], [
        ("=", LONG_A, [True, False, False]),
        ("=", "prefix_" + "a" * 200, [False, False, False]),
        ("=", LONG_B, [False, True, False]),
    ]),
    ("not_equals_preserves_nulls (tests.rs:776-785)", ["alpha", None, "beta", "alpha"], [
        ("!=", "alpha", [False, None, True, False]),
    ]),
    ("shared_prefix_shorter_needle_lt (tests.rs:808-822)", ["hello_world", "hello_rust"], [
        ("<", "hell", [False, False]),
        ("<=", "hell", [False, False]),
    ]),
]

# fingerprinted arrays (hint SubstringSearch): tests.rs:106-140
FINGERPRINT_CASES = [
    ("fingerprint_skips_impossible_substring (tests.rs:106-140)", ["alpha", "ALP", "beta", "gamma"], [
        ("like", "%zzz%", [False, False, False, False]),
        ("like", "%alp%", [True, False, False, False]),
    ]),
]

# no fingerprints -> LIKE falls back to arrow semantics (tests.rs:824-851)
LIKE_FALLBACK_CASES = [
    ("like_fallback (tests.rs:824-851)", ["Alpha", "alphabet", "beta", None, "ALPHA"], [
        ("like", "Al%", [True, False, False, None, False]),
        ("not like", "Al%", [False, True, True, None, True]),
    ]),
]

# PrefixKey bytes and shared prefixes: tests.rs:176-393
PREFIX_KEY_CASES = [
    (["hello", "world", "test"], b"", [b"hello\0\0", b"world\0\0", b"test\0\0\0"]),
    (["hello_world", "hello_rust", "hello_test", "hello_code"], b"hello_", [b"world\0\0", b"rust\0\0\0", b"test\0\0\0", b"code\0\0\0"]),
    (["abc", "abcde", "abcdef", "abcdefg"], b"abc", [b"\0" * 7, b"de\0\0\0\0\0", b"def\0\0\0\0", b"defg\0\0\0"]),
    (["data", "database", "data_entry", "data_", "datatype"], b"data",
     [b"\0" * 7, b"base\0\0\0", b"_entry\0", b"_\0\0\0\0\0\0", b"type\0\0\0"]),
    (["identical", "identical", "identical"], b"identical", [b"\0" * 7]),
    (["hello", "hello_world", "hello_test"], b"hello", [b"\0" * 7, b"_world\0", b"_test\0\0"]),
    (["", "hello", "hello_world"], b"", [b"\0" * 7, b"hello\0\0", b"hello_w"]),
]

# src/datafusion/src/reader/utils/boolean_selection.rs:233-256
AND_THEN_CASE = ("001011010101", "001101", "000001010001")

# README.md:43-88 and src/core/README.md:17-104
QUICK_START = {
    "values": [10, 11, 12, 13, 14, 15],
    "selection": [True, False, True, False, True, False],
    "filtered": [10, 12, 14],
    "gt12": [False, False, False, True, True, True],
    "strings": ["apple", "banana", None, "apple", "cherry"],
    "string_selection": [True, True, False, True, True],
    "eq_apple_selected": [True, False, True, False],
}

# Multi-column OR over cached columns (src/datafusion/src/cache/mod.rs:433-639): (columns, conjuncts, rows that match)
MULTI_COLUMN_OR_CASES = [
    # evaluate_or_on_cached_columns: a = 3 OR b = 20
    ([("int32", [1, 2, 3, 4]), ("int32", [10, 20, 30, 40])], [("=", 3), ("=", 20)], [1, 2]),
    # evaluate_three_column_or: a = 2 OR b = 40 OR c = 600
    ([("int32", [1, 2, 3, 4, 5, 6, 7, 8]), ("int32", [10, 20, 30, 40, 50, 60, 70, 80]), ("int32", [100, 200, 300, 400, 500, 600, 700, 800])],
     [("=", 2), ("=", 40), ("=", 600)], [1, 3, 5]),
    # evaluate_string_column_or: name = 'Bob' OR city = 'Tokyo' (Utf8View columns)
    ([("string_view", ["Alice", "Bob", "Charlie", "David", "Eve", "Frank", "Grace", "Henry"]),
      ("string_view", ["New York", "London", "Paris", "Tokyo", "Berlin", "Sydney", "Madrid", "Rome"])], [("=", "Bob"), ("=", "Tokyo")], [1, 3]),
]
