"""Log-line helper the reference Cython imports (Utils/seconds_to_biggest_unit.py:11); restated, TEST INFRASTRUCTURE."""


def seconds_to_biggest_unit(time_in_seconds, data_array=None):
    value, unit = time_in_seconds, "sec"
    for limit, name in ((60, "min"), (60, "hour"), (24, "day")):
        if value < limit:
            break
        value /= limit
        if data_array is not None:
            data_array = data_array / limit
        unit = name
    return (value, unit) if data_array is None else (value, unit, data_array)
