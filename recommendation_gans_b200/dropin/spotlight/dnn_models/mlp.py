"""Multi-layer perceptron scorer for `representation=` (reference: spotlight/dnn_models/mlp.py:5-46).

`MLP(layers, num_users, num_items, output_dim=1, embedding_dim=32)`: user and item embeddings are concatenated and
pushed through Linear -> LeakyReLU(0.1) -> Dropout(0.5) blocks of the widths in `layers` (layers[0] must be
2 * embedding_dim) and a final Linear(layers[-1], 1) + sigmoid; output shape [n, 1].  Same sub-module names and
state-dict keys as the reference.  ImplicitFactorizationModel trains it on the generic torch-autograd step, with the
negative pairs still drawn on the device and the top-k evaluation still done by the CUDA ranking kernel."""
import torch
import torch.nn as nn


def _init_linear(module):
    if type(module) == nn.Linear:
        nn.init.xavier_uniform_(module.weight)
        module.bias.data.fill_(0.01)


def _tower(widths):
    blocks = nn.ModuleList()
    for fan_in, fan_out in zip(widths[:-1], widths[1:]):
        blocks.append(nn.Linear(fan_in, fan_out))
        blocks.append(nn.LeakyReLU(0.1, inplace=True))
        blocks.append(nn.Dropout(0.5))
    return blocks


class MLP(nn.Module):

    def __init__(self, layers, num_users, num_items, output_dim=1, embedding_dim=32):
        super(MLP, self).__init__()
        self.num_users, self.num_items, self.latent_dim = num_users, num_items, embedding_dim
        self.embedding_user = nn.Embedding(num_embeddings=num_users, embedding_dim=embedding_dim)
        self.embedding_item = nn.Embedding(num_embeddings=num_items, embedding_dim=embedding_dim)
        self.layers = _tower(layers)
        self.layers.append(nn.Linear(layers[-1], out_features=1))
        self.logistic = nn.Sigmoid()
        self.apply(self.init_weights)

    def forward(self, user_indices, item_indices):
        vector = torch.cat([self.embedding_user(user_indices), self.embedding_item(item_indices)], dim=-1)
        for layer in self.layers:
            vector = layer(vector)
        return self.logistic(vector)

    def init_weights(self, m):
        _init_linear(m)
