"""Experiment bookkeeping used by fit() (reference: utils/storage_utils.py:6-84): per-epoch CSV
statistics and checkpoint-key helpers.  Host-side file I/O, same formats as the reference."""
import csv
import os
import pickle

import torch


def dict_load(model_path, parallel=False):
    """Load a `{'network': state_dict}` checkpoint and strip the 6- (or 13-) character key prefix."""
    skip = 13 if parallel else 6
    blob = torch.load(model_path, map_location=None if torch.cuda.is_available() else 'cpu')
    return {key[skip:]: value for key, value in blob['network'].items()}


def save_to_stats_pkl_file(experiment_log_filepath, filename, stats_dict):
    with open(os.path.join(experiment_log_filepath, filename) + '.pkl', 'wb') as fh:
        pickle.dump(stats_dict, fh)


def load_from_stats_pkl_file(experiment_log_filepath, filename):
    with open(os.path.join(experiment_log_filepath, filename) + '.pkl', 'rb') as fh:
        return pickle.load(fh)


def save_statistics(experiment_log_dir, filename, stats_dict, current_epoch, continue_from_mode=False,
                    save_full_dict=False):
    """Append (or start) `filename` with one row per epoch; header = the dict's keys."""
    path = os.path.join(experiment_log_dir, filename)
    columns = list(stats_dict.values())
    with open(path, 'a' if continue_from_mode else 'w') as fh:
        writer = csv.writer(fh)
        if not continue_from_mode:
            writer.writerow(list(stats_dict.keys()))
        if save_full_dict:
            for idx in range(len(columns[0])):
                writer.writerow([col[idx] for col in columns])
        else:
            writer.writerow([col[current_epoch] for col in columns])
    return path


def load_statistics(experiment_log_dir, filename):
    with open(os.path.join(experiment_log_dir, filename), 'r+') as fh:
        lines = fh.readlines()
    keys = lines[0].split(",")
    stats = {key: [] for key in keys}
    for line in lines[1:]:
        for idx, value in enumerate(line.split(",")):
            stats[keys[idx]].append(value)
    return stats
