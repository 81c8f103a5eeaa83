"""Neural matrix factorisation scorer for `representation=` (reference: spotlight/dnn_models/neuMF.py:7-61).

`NeuMF(mlp_layers, num_users, num_items, mf_embedding_dim=25, mlp_embedding_dim=32)`: an MLP tower over the concatenated
MLP embeddings and a GMF half (element-wise product of the MF embeddings), concatenated into Linear(., 1) + sigmoid;
output shape [n, 1].  Same sub-module names and state-dict keys as the reference."""
import torch
import torch.nn as nn

from spotlight.dnn_models.mlp import _init_linear, _tower


class NeuMF(nn.Module):

    def __init__(self, mlp_layers, num_users, num_items, mf_embedding_dim=25, mlp_embedding_dim=32):
        super(NeuMF, self).__init__()
        self.num_users, self.num_items = num_users, num_items
        self.latent_dim_mf, self.latent_dim_mlp = mf_embedding_dim, mlp_embedding_dim
        self.embedding_user_mlp = nn.Embedding(num_embeddings=num_users, embedding_dim=mlp_embedding_dim)
        self.embedding_item_mlp = nn.Embedding(num_embeddings=num_items, embedding_dim=mlp_embedding_dim)
        self.embedding_user_mf = nn.Embedding(num_embeddings=num_users, embedding_dim=mf_embedding_dim)
        self.embedding_item_mf = nn.Embedding(num_embeddings=num_items, embedding_dim=mf_embedding_dim)
        self.layers = _tower(mlp_layers)
        self.affine_output = nn.Linear(mlp_layers[-1] + mf_embedding_dim, out_features=1)
        self.logistic = nn.Sigmoid()
        self.apply(self.init_weights)

    def forward(self, user_indices, item_indices):
        mlp_vector = torch.cat([self.embedding_user_mlp(user_indices), self.embedding_item_mlp(item_indices)], dim=-1)
        mf_vector = torch.mul(self.embedding_user_mf(user_indices), self.embedding_item_mf(item_indices))
        for layer in self.layers:
            mlp_vector = layer(mlp_vector)
        return self.logistic(self.affine_output(torch.cat([mlp_vector, mf_vector], dim=-1)))

    def init_weights(self, m):
        _init_linear(m)
