"""Command-line flags of the training scripts (reference: utils/arg_extractor.py:15-78).

Same flag names, types and defaults, so `mf_spotlight.py` and the cluster scripts' command lines keep working."""
import argparse


def str2bool(v):
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


# (flag, type, default, help); str2bool flags also accept a bare `--flag` where the reference does (nargs='?')
_FLAGS = (
    # all experiments
    ('--use_gpu', str2bool, False, 'run on the GPU (this build has no CPU path: the flag is accepted either way)'),
    ('--l2_regularizer', float, 1e-5, 'l2 normalization constant'),
    ('--on_cluster', str2bool, False, 'read the datasets from the cluster scratch disk'),
    # learning to rank
    ('--model', str, 'mf', 'mf/mlp/neuMF'),
    ('--dataset', str, '100K', '100K/1M/10M/20M'),
    ('--experiment_name', str, 'matrix_model', 'name of the resulting experiment'),
    ('--precision_recall', str2bool, True, 'compute precision/recall at k'),
    ('--map_recall', str2bool, True, 'compute mean average precision / recall at k'),
    ('--rmse', str2bool, True, 'compute the root mean square error'),
    ('--mf_embedding_dim', int, 50, 'latent dimensions of the matrix factorization model'),
    ('--mlp_embedding_dim', int, 16, 'latent dimensions of the mlp embeddings'),
    ('--training_epochs', int, 50, 'training epochs'),
    ('--batch_size', int, 256, 'minibatch size'),
    ('--learning_rate', float, 1e-3, 'learning rate'),
    ('--optim', str, 'adam', 'adam/sgd/rms'),
    ('--k', int, 3, 'cut-off of precision@k / recall@k'),
    ('--neg_examples', int, 5, 'negative examples per positive'),
    # slate generation
    ('--optim_gan', str, 'rms', 'adam/sgd/rms'),
    ('--gan_embedding_dim', int, 5, 'latent dimensions of the GAN embeddings'),
    ('--gan_hidden_layer', int, 10, 'hidden layer width of the GAN'),
    ('--loss', str, 'bce', 'bce/mse'),
    ('--slate_size', int, 3, 'size of the generated slate'),
)


def get_args():
    """Arguments extracted from the command line (an argparse.Namespace)."""
    parser = argparse.ArgumentParser(description='implicit-feedback recommender training / evaluation')
    for flag, kind, default, text in _FLAGS:
        extra = {'nargs': '?'} if flag == '--use_gpu' else {}
        parser.add_argument(flag, type=kind, default=default, help=text, **extra)
    return parser.parse_args()
