"""speechbrain.nnet.containers stand-in.

``Sequential`` is a ModuleDict whose ``append`` (a) de-duplicates layer names as name, name_0,
name_1, ... and (b) instantiates layer *classes* that take ``input_shape`` with the shape a dummy
forward through the layers so far produces.  That naming rule yields the ``linear`` / ``linear_0``
state_dict keys of VanillaNN (reference VanillaNN.py:181-196).
"""
import inspect

import torch


class Sequential(torch.nn.ModuleDict):
    def __init__(self, *layers, input_shape=None, **named_layers):
        super().__init__()
        if not layers and input_shape is None and not named_layers:
            raise ValueError("Must pass either layers or input shape")
        self.length_layers = []
        self.input_shape = input_shape
        if input_shape and None in input_shape:
            self.input_shape = list(input_shape)
            for i, dim in enumerate(self.input_shape):
                if i == 0 and dim is None:
                    dim = 1
                self.input_shape[i] = dim or 256
        for layer in layers:
            self.append(layer)
        for name, layer in named_layers.items():
            self.append(layer, layer_name=name)

    def append(self, layer, *args, layer_name=None, **kwargs):
        if layer_name is None:
            layer_name = str(len(self))
        elif layer_name in self:
            index = 0
            while f"{layer_name}_{index}" in self:
                index += 1
            layer_name = f"{layer_name}_{index}"
        if self.input_shape:
            argspec = inspect.getfullargspec(layer)
            if "input_shape" in argspec.args + argspec.kwonlyargs:
                input_shape = self.get_output_shape()
                layer = layer(*args, input_shape=input_shape, **kwargs)
        try:
            self.add_module(layer_name, layer)
        except TypeError:
            raise ValueError("Must pass `input_shape` at initialization and use modules that take `input_shape`")

    def get_output_shape(self):
        with torch.no_grad():
            dummy_input = torch.zeros(self.input_shape)
            dummy_output = self(dummy_input)
        return dummy_output.shape

    def forward(self, x):
        for layer in self.values():
            x = layer(x)
            if isinstance(x, tuple):
                x = x[0]
        return x


class ModuleList(torch.nn.Module):
    def __init__(self, *layers):
        super().__init__()
        self.layers = torch.nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
            if isinstance(x, tuple):
                x = x[0]
        return x

    def append(self, module):
        self.layers.append(module)

    def extend(self, modules):
        self.layers.extend(modules)

    def insert(self, index, module):
        self.layers.insert(index, module)
