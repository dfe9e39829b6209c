"""Special vocabulary ids (reference data_utils.py:13-15).  The word filtering of data_utils.py:17-33 lives in
scoring.py; vocabulary files (data_utils.py:35-62) are out of scope."""

PAD_ID = 0
GO_ID = 1
EOS_ID = 2
