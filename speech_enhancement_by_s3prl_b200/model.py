"""Drop-in mask heads (reference model.py).  Same class names, constructor keywords and
``forward(features=, linears=) -> (predicted, dict)`` contract, because the reference
builds them with ``eval(args.downstream)(input_size=, output_size=, **all_cli_args)``
(run_downstream.py:208-210) and calls them by keyword (runner.py:453, 569; sampler.py:71).

``Linear`` / ``LinearResidual`` run on the library's head kernel (CMVN folded into the
operand load, bias + activation in the epilogue).  ``LSTM`` / ``Residual`` keep the stock
cuDNN ``nn.LSTM`` body -- recurrent, outside the named path -- and use the
head kernel for their projection layer.
"""
import torch
import torch.nn as nn

from . import ops

_ACTS = ("Identity", "ReLU", "Sigmoid")


def _check_act(name):
    if name not in _ACTS:
        raise ValueError(f"activation {name!r} is not supported by the fused head (choose from {_ACTS})")
    return name


class Linear(nn.Module):
    """model.py:8-17 -- predicted = act(W features + b)."""

    def __init__(self, input_dim=None, output_dim=None, activation="ReLU", input_size=None, output_size=None,
                 precision=0, **kwargs):
        super().__init__()
        input_dim = input_size if input_dim is None else input_dim
        output_dim = output_size if output_dim is None else output_dim
        self.linear = nn.Linear(input_dim, output_dim)
        self.activation = _check_act(activation)
        self.precision = precision

    def forward(self, features, **kwargs):
        predicted = ops.linear_head(features, self.linear.weight, self.linear.bias, self.activation, precision=self.precision)
        return predicted, {}


class LinearResidual(nn.Module):
    """model.py:20-34 -- CMVN over time -> Linear -> act -> predicted = linears * offset."""

    def __init__(self, input_size=201, output_size=201, activation="Sigmoid", cmvn=True, eps=1e-6, precision=0, **kwargs):
        super().__init__()
        self.linear = nn.Linear(input_size, output_size)
        self.activation = _check_act(activation)
        self.cmvn = cmvn
        self.eps = eps
        self.precision = precision

    def forward(self, features, linears, **kwargs):
        mean = std = None
        if self.cmvn:
            mean, std = ops.cmvn_stats(features)
        offset = ops.linear_head(features, self.linear.weight, self.linear.bias, self.activation, mean, std, self.eps,
                                 precision=self.precision)
        predicted = linears * offset
        return predicted, {"offset": offset}


def _init_lstm_like(module):
    for name, param in module.named_parameters():
        if "weight_ih" in name or "scaling_layer.0.weight" in name:
            nn.init.xavier_uniform_(param.data)
        elif "weight_hh" in name:
            nn.init.orthogonal_(param.data)
        elif "bias" in name:
            nn.init.constant_(param.data, 0)


def _project(hidden, lin, activation, precision):
    """Projection after the recurrent body: the library head, forward and backward (weight, bias and -- while the LSTM
    trains -- input gradients all come from the head kernels)."""
    return ops.linear_head(hidden.contiguous(), lin.weight, lin.bias, activation, precision=precision)


class LSTM(nn.Module):
    """model.py:37-60 -- LSTM -> Linear+act = log_predicted; predicted = exp(log_predicted)."""

    def __init__(self, input_size=201, output_size=201, hidden_size=201, num_layers=3, bidirectional=False,
                 activation="Identity", precision=0, **kwargs):
        super().__init__()
        self.lstm = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, batch_first=True,
                            bidirectional=bidirectional)
        self.scaling_layer = nn.Sequential(nn.Linear((2 if bidirectional else 1) * hidden_size, output_size), nn.Identity())
        self.activation = _check_act(activation)
        self.bidirectional = bidirectional
        self.precision = precision
        _init_lstm_like(self)

    def forward(self, features, **kwargs):
        hidden, _ = self.lstm(features)
        log_predicted = _project(hidden, self.scaling_layer[0], self.activation, self.precision)
        return log_predicted.exp(), {"log_predicted": log_predicted}


class Residual(nn.Module):
    """model.py:63-91 -- LSTM -> [CMVN] -> Linear+act = offset; predicted = linears * offset."""

    def __init__(self, input_size=201, output_size=201, hidden_size=201, num_layers=3, bidirectional=False,
                 activation="Sigmoid", cmvn=False, eps=1e-6, precision=0, **kwargs):
        super().__init__()
        self.lstm = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, batch_first=True,
                            bidirectional=bidirectional)
        self.scaling_layer = nn.Sequential(nn.Linear((2 if bidirectional else 1) * hidden_size, output_size), nn.Identity())
        self.activation = _check_act(activation)
        self.bidirectional = bidirectional
        self.cmvn = cmvn
        self.eps = eps
        self.precision = precision
        _init_lstm_like(self)

    def forward(self, features, linears, **kwargs):
        hidden, _ = self.lstm(features)
        if self.cmvn:
            # the LSTM output needs a gradient, so this CMVN stays in autograd (model.py:88)
            hidden = (hidden - hidden.mean(dim=1, keepdim=True)) / (hidden.std(dim=1, keepdim=True) + self.eps)
        offset = _project(hidden, self.scaling_layer[0], self.activation, self.precision)
        predicted = linears * offset
        return predicted, {"offset": offset}
