"""Model hand-off in the reference's on-disk format.

`DataIO(folder_path).save_data(file_name, dict)` / `.load_data(file_name)` read and write the zip archives of
Base/DataIO.py:105-253: one member per attribute -- scipy sparse -> `<name>.npz` (scipy.sparse.save_npz), ndarray ->
`<name>.npy` (allow_pickle=False), pandas DataFrame -> `<name>.csv`, JSON-serialisable values -> `<name>.json`, nested
dict that JSON refuses -> `<name>.zip` (recursively), plus the manifest `.DataIO_attribute_to_file_name.json` -- so a
model saved by either implementation loads in the other.  Written against the format, not the reference's code: members
are produced in memory (no `.temp` folder on disk, Base/DataIO.py:59-79), which also makes concurrent savers safe.
Host-only; no device work."""
import io
import json
import os
import zipfile

import numpy as np
import scipy.sparse as sps

_MANIFEST = ".DataIO_attribute_to_file_name"
_OLD_MANIFEST = "__DataIO_attribute_to_file_name"  # Base/DataIO.py:211: archives of older reference versions


def _json_default(o):
    """Base/DataIO.py:17-31: numpy scalars are not JSON-serialisable by themselves."""
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, np.bool_):
        return bool(o)
    raise TypeError("json_not_serializable_handler: object '{}' is not serializable.".format(type(o)))


def _str_keys(d):
    """Base/DataIO.py:81-102: JSON objects have string keys only."""
    if all(isinstance(k, str) for k in d):
        return d
    out = {str(k): v for k, v in d.items()}
    assert len(out) == len(d), "DataIO: Transforming dictionary keys into strings altered its content. Duplicate keys may have been produced."
    return out


def _encode(data_dict_to_save):
    """dict -> bytes of one archive."""
    members, manifest = {}, {}
    for name, value in data_dict_to_save.items():
        buf = io.BytesIO()
        if type(value).__name__ == "DataFrame" and hasattr(value, "to_csv"):
            members[name + ".csv"] = value.to_csv(index=False).encode()
            manifest[name] = name + ".csv"
        elif sps.issparse(value):
            sps.save_npz(buf, value)
            members[name + ".npz"] = buf.getvalue()
            manifest[name] = name + ".npz"
        elif isinstance(value, np.ndarray):
            np.save(buf, value, allow_pickle=False)
            members[name + ".npy"] = buf.getvalue()
            manifest[name] = name + ".npy"
        else:
            try:
                text = json.dumps(_str_keys(value) if isinstance(value, dict) else value, default=_json_default)
                members[name + ".json"] = text.encode()
                manifest[name] = name + ".json"
            except TypeError:
                if not isinstance(value, dict):
                    raise TypeError("Type not recognized for attribute: {}".format(name))
                members[name + ".zip"] = _encode(value)
                manifest[name] = name + ".zip"
    members[_MANIFEST + ".json"] = json.dumps(manifest).encode()
    out = io.BytesIO()
    with zipfile.ZipFile(out, "w", compression=zipfile.ZIP_DEFLATED) as z:
        for member, payload in members.items():
            z.writestr(member, payload)
    return out.getvalue()


def _decode(raw):
    with zipfile.ZipFile(io.BytesIO(raw)) as z:
        names = set(z.namelist())
        manifest_name = _MANIFEST + ".json" if _MANIFEST + ".json" in names else _OLD_MANIFEST + ".json"
        manifest = json.loads(z.read(manifest_name).decode())
        out = {}
        for attrib, member in manifest.items():
            payload = z.read(member)
            kind = member.split(".")[-1]
            if kind == "csv":
                import pandas as pd
                out[attrib] = pd.read_csv(io.BytesIO(payload), index_col=False)
            elif kind == "npz":
                out[attrib] = sps.load_npz(io.BytesIO(payload))
            elif kind == "npy":
                out[attrib] = np.load(io.BytesIO(payload), allow_pickle=False)
            elif kind == "zip":
                out[attrib] = _decode(payload)
            elif kind == "json":
                out[attrib] = json.loads(payload.decode())
            else:
                raise Exception("Attribute type not recognized for: '{}' of class: '{}'".format(member, kind))
        return out


class DataIO(object):
    def __init__(self, folder_path):
        self.folder_path = folder_path

    @staticmethod
    def _zip_name(file_name):
        return file_name if file_name[-4:] == ".zip" else file_name + ".zip"

    def save_data(self, file_name, data_dict_to_save):
        if not os.path.exists(self.folder_path):
            os.makedirs(self.folder_path)
        path = self.folder_path + self._zip_name(file_name)
        tmp = path + ".partial.%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(_encode(data_dict_to_save))
        os.replace(tmp, path)  # readers never see a half-written archive

    def load_data(self, file_name):
        with open(self.folder_path + self._zip_name(file_name), "rb") as f:
            return _decode(f.read())
