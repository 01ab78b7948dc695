"""reference modems_codecs/string_ops.py:6-15"""


def check_boolean(input_string):
	return input_string.lower() in ("yes", "true", "1")
