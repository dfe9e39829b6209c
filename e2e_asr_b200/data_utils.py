"""Special vocabulary ids (reference data_utils.py:13-15). Text post-processing
(data_utils.py:17-62) is out of scope."""

PAD_ID = 0
GO_ID = 1
EOS_ID = 2
