"""Small helpers used by the grid-search wrappers (reference: helper_funcs.py:1-30)."""


def get_secs_mins_hours_from_secs(total_secs):
    """(hours, minutes, seconds) of a duration; integer division as the Python-2 reference did."""
    total_secs = int(total_secs)
    return total_secs // 3600, (total_secs % 3600) // 60, total_secs % 60


def get_friendly_label_name(col):
    """'..._happiness_label' -> 'happiness' etc. (reference: helper_funcs.py:18-30)."""
    if col is None:
        return ""
    if not isinstance(col, str):
        return str(col)
    low = col.lower()
    for key in ('happiness', 'calmness', 'health'):
        if key in low:
            return key
    return ""
