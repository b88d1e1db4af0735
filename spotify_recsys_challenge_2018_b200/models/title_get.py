"""get_model(conf): host mirror of the reference's models/title_get.py:10-26.

Only Char_CNN is reachable in the reference (Char_LSTM's branch calls an undefined name and reads a
config field main.py never sets) -- asking for it fails loudly here too."""
from .title_models.Char_CNN import Char_CNN


def get_model(conf):
    model = conf.char_model
    assert model in ["Char_CNN", "Char_LSTM"]                     # title_get.py:12
    if model == "Char_CNN":
        conv_layers = [[conf.filter_num, fs, -1] for fs in conf.filter_size]      # title_get.py:14-20
        return Char_CNN(conf, conv_layers)
    raise NotImplementedError("Char_LSTM is unreachable in the reference (title_get.py:26 calls an undefined name)")
