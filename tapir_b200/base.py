"""Command-line and filesystem helpers of tapir's driver, mirrored from tapir/base.py:19-123 so that
bin/tapir_compute.py keeps its argument types, glob rules and output-directory rule.  Pure host glue:
nothing here computes."""
import argparse
import errno
import glob
import os


class FullPaths(argparse.Action):
    """argparse action: store the absolute, user-expanded path (tapir/base.py:19-22)."""

    def __call__(self, parser, namespace, values, option_string=None):
        setattr(namespace, self.dest, to_full_paths(values))


def to_full_paths(string):
    return os.path.abspath(os.path.expanduser(string))


def is_dir(dirname):
    """argparse type: the argument must name an existing directory (tapir/base.py:27-33)."""
    if os.path.isdir(dirname):
        return dirname
    raise argparse.ArgumentTypeError("{0} is not a directory".format(dirname))


def mkdir(path):
    try:
        os.mkdir(path)
    except OSError:
        pass


def create_unique_dir(path, limit=100):
    """Use `path` when it is empty, else create path.1, path.2, ... (tapir/base.py:41-66)."""
    if not os.listdir(path):
        return path
    candidate = path
    for count in range(1, limit + 1):
        try:
            os.mkdir(candidate)
            return candidate
        except OSError as err:
            if err.errno != errno.EEXIST:
                raise
            candidate = "{0}.{1}".format(path, count)
            if count == limit:
                break
    raise Exception("could not uniquely create directory {0}: limit `{1}` reached".format(path, limit))


def get_output_type(name):
    ext = os.path.splitext(name)[1].lstrip(".").lower()
    assert ext in ("pdf", "png", "tiff", "jpeg", "jpg"), "Filetype must be one of pdf, png, tiff, or jpeg"
    return ext


def _split(string, convert, what, msg):
    try:
        return [convert(item) for item in string.split(",")]
    except Exception as err:
        raise argparse.ArgumentTypeError(msg.format(what, err))


def get_list_from_ints(string, name="time"):
    """'37,93' -> [37, 93] (tapir/base.py:75-82)."""
    return _split(string, int, name, "Cannot convert {0} to a list of integers: {1}")


def get_strings_from_items(string, name="locus"):
    return _split(string, str, name, "Cannot convert {0} to a list of loci: {1}")


def get_list_from_ranges(string):
    """'32-42,88-98' -> [[32, 42], [88, 98]]; extra fields of a span are kept, as in tapir/base.py:93-100."""
    try:
        return [[int(bound) for bound in span.split("-")] for span in string.split(",")]
    except Exception as err:
        raise argparse.ArgumentTypeError("Cannot convert spans to a list of integers: {0}".format(err))


def get_files(d, extension):
    """All files of `d` matching a glob or a comma-separated list of globs, in glob order
    (tapir/base.py:102-114); IOError when there are none."""
    patterns = extension.strip(" ").split(",") if "," in extension else [extension]
    files = []
    for pattern in patterns:
        files.extend(glob.glob(os.path.join(d, pattern)))
    if not files:
        raise IOError("There appear to be no files of {0} type in {1}".format(patterns, d))
    return files


def parse_subset_map_file(filename):
    """alignment<TAB>start<TAB>end lines -> (alignment, [start, end]) pairs (tapir/base.py:116-123)."""
    with open(filename) as handle:
        for line in handle:
            line = line.strip()
            if line:
                align, start, end = line.split("\t", 2)
                yield align, [int(start), int(end)]
