"""Label tables: id <-> character maps.  Same contents and names as the reference's ``data/label_sets.py:2-14``
(blank '_' at index 0, apostrophe, the alphabet, space last), built programmatically."""
import string


def _with_ctc_symbols(symbols):
    return ["_"] + list(symbols) + [" "]          # CTC blank first (blank index 0), space last


english_labels = _with_ctc_symbols(["'"] + list(string.ascii_uppercase))
english_lowercase_labels = _with_ctc_symbols(["'"] + list(string.ascii_lowercase))
# Hebrew: the 22 letters in alphabet order (final forms excluded), then the five final forms
_HEB_FINALS = [0x05DF, 0x05E3, 0x05E5, 0x05DD, 0x05DA]
hebrew_labels = _with_ctc_symbols([chr(c) for c in range(0x05D0, 0x05EB) if c not in (0x05DA, 0x05DD, 0x05DF, 0x05E3, 0x05E5)]
                                  + [chr(c) for c in _HEB_FINALS])

labels_map = {"english": english_labels, "hebrew": hebrew_labels, "english_lowercase": english_lowercase_labels}
