"""Implicit-feedback factorisation model (reference: implicit.py:30-471), B200-native.

`ImplicitFactorizationModel` keeps the reference's constructor, attributes, side effects
(experiment directories, configuration.json, summary.csv, test_summary.json, best_model) and
error behaviour, but its inner loops run as CUDA kernels through the mfb200 C ABI:

  fit()                  one planner + fused step pipeline per epoch (mfb_train_steps), validation
                         losses (mfb_loss_steps); per-step losses come back once per epoch
  run_train_iteration()  a single fused step (same kernels), returns the batch loss tensor
  predict()              fused gather/dot/sigmoid (mfb_predict_pairs / mfb_predict_user)
  test()                 full-catalog top-k with fused mask (mfb_topk) + hit counting

Negative pairs are drawn exactly as the reference does -- `random.choices(neg_examples, k)` on
Python's global MT19937 stream (implicit.py:352,370) -- but on the device; the `random` module's
state is advanced to where the reference would leave it.  There is no CPU path: `use_cuda=False`
still runs on the current CUDA device.
"""
import copy
import json
import logging
import os
import random

import numpy as np
import torch
import torch.optim as optim

from recommendation_gans_b200.engine import MFEngine
from spotlight.evaluation import precision_recall_score, rmse_score
from spotlight.factorization._components import _predict_process_ids
from spotlight.factorization.representations import BilinearNet
from spotlight.helpers import _repr_model
from spotlight.torch_utils import minibatch, set_seed, shuffle
from utils.storage_utils import save_statistics

logging.basicConfig(format='%(message)s', level=logging.INFO)

_LOSSES = ('pointwise', 'bpr', 'hinge', 'adaptive_hinge')


class ImplicitFactorizationModel(object):

    def __init__(self, loss='pointwise', embedding_dim=32, n_iter=10, batch_size=256, l2=0.0,
                 experiment_name='Implicit_Feedback', learning_rate=1e-2, optimizer_func=None,
                 use_cuda=False, representation=None, sparse=False, model_name='mf',
                 random_state=None, neg_examples=None, num_negative_samples=3):
        self.exeriment_name = experiment_name
        self.experiment_folder = os.path.abspath(os.path.join("experiments_results", experiment_name))
        self.experiment_logs = os.path.join(self.experiment_folder, "result_outputs")
        self.experiment_saved_models = os.path.join(self.experiment_folder, "saved_models")
        self.starting_epoch = 0
        for folder in (self.experiment_folder, self.experiment_logs, self.experiment_saved_models):
            os.makedirs(folder, exist_ok=True)

        assert loss in _LOSSES

        self._loss = loss
        self._embedding_dim = embedding_dim
        self._n_iter = n_iter
        self._learning_rate = learning_rate
        self._batch_size = batch_size
        self._l2 = l2
        self._use_cuda = use_cuda
        self._representation = representation
        self._sparse = sparse
        self._optimizer_func = optimizer_func
        self._random_state = random_state or np.random.RandomState()
        self._num_negative_samples = num_negative_samples
        self.neg_examples = neg_examples

        self._num_users = None
        self._num_items = None
        self._net = None
        self._optimizer = None
        self._loss_func = None
        self._loss_kind = None
        self._engine_ = None
        self._generic = False
        self._neg_pop = None
        self.best_model = None
        self.best_validation = None
        self.model_name = model_name
        self.best_epoch = -1

        if not torch.cuda.is_available():
            raise RuntimeError('ImplicitFactorizationModel (mfb200) needs a CUDA device; there is no CPU path.')
        if not use_cuda:
            logging.info("mfb200: use_cuda=False requested, but this build only has a CUDA path; "
                         "running on cuda:%d", torch.cuda.current_device())
        # implicit.py:146-147 (consumes one draw of the model's RandomState)
        set_seed(self._random_state.randint(-10 ** 8, 10 ** 8), cuda=True)

    def __repr__(self):
        return _repr_model(self)

    @property
    def _initialized(self):
        return self._net is not None

    # -------------------------------------------------------------------------------------------
    def _bind(self, net):
        """BilinearNet runs on the fused CUDA step; any other torch module (`representation=` MLP, NeuMF, ...,
        implicit.py:169-180) trains on the generic torch-autograd step -- with the negative pairs still drawn on the
        device from Python's `random` stream and the evaluation still ranked by the CUDA top-k kernel."""
        self._generic = not isinstance(net, BilinearNet)
        if not self._generic and (self._sparse or any(getattr(l, 'sparse', False) for l in
                                  (net.user_embeddings, net.item_embeddings, net.user_biases, net.item_biases))):
            raise NotImplementedError('mfb200: sparse=True embeddings are not supported '
                                      '(the reference script always uses sparse=False)')
        return net.cuda()

    def set_users(self, _num_users, _num_items):
        self._num_users = _num_users
        self._num_items = _num_items
        self._net = self._bind(self._representation)

    def _initialize(self, interactions):
        (self._num_users, self._num_items) = (interactions.num_users, interactions.num_items)
        if self._net is not None:
            # A later fit() call: keep training the network the previous call left (best_model, implicit.py:338) with
            # a fresh optimiser and engine bound to ITS parameters.  (The reference skips _initialize here and keeps
            # stepping an optimiser that still points at the pre-deepcopy parameters, so its second fit() trains
            # nothing; resuming from the kept weights is what its docstring promises.)
            self._net = self._bind(self._net)
        elif self._representation is not None:
            self._net = self._bind(self._representation)
        else:
            self._net = self._bind(BilinearNet(self._num_users, self._num_items, self._embedding_dim,
                                               sparse=self._sparse))
        if self._optimizer_func is None:
            self._optimizer = optim.Adam(self._net.parameters(), weight_decay=self._l2, lr=self._learning_rate)
        else:
            self._optimizer = self._optimizer_func(self._net.parameters(), weight_decay=self._l2,
                                                   lr=self._learning_rate)
        # the reference wires only 'pointwise' and 'hinge' by name; every other accepted name
        # (including 'bpr') trains with adaptive_hinge_loss (implicit.py:194-199)
        from spotlight import losses as _losses
        if self._loss == 'pointwise':
            self._loss_kind, self._loss_func = 'pointwise', _losses.pointwise_loss
        elif self._loss == 'hinge':
            self._loss_kind, self._loss_func = 'hinge', _losses.hinge_loss
        else:
            self._loss_kind, self._loss_func = 'adaptive_hinge', _losses.adaptive_hinge_loss
        if self._generic:
            self._engine_ = None
        else:
            self._engine_ = MFEngine(self._net, self._optimizer)
            self._net._attach_engine(self._engine_)
        self.configuration = {
            'num_users': self._num_users, 'num_items': self._num_items, 'weight_decay': self._l2,
            'lr': self._learning_rate, 'embedding_dim': self._embedding_dim,
            'batch_size': self._batch_size, 'epochs': self._n_iter}
        with open(os.path.join(self.experiment_logs, 'configuration.json'), 'w') as fp:
            json.dump(self.configuration, fp)

    def _check_input(self, user_ids, item_ids, allow_items_none=False):
        user_id_max = user_ids if isinstance(user_ids, int) else user_ids.max()
        if user_id_max >= self._num_users:
            raise ValueError('Maximum user id greater than number of users in model.')
        if allow_items_none and item_ids is None:
            return
        item_id_max = item_ids if isinstance(item_ids, int) else item_ids.max()
        if item_id_max >= self._num_items:
            raise ValueError('Maximum item id greater than number of items in model.')

    # -------------------------------------------------------------------------------------------
    def _negative_population(self):
        """neg_examples (list of (user, item) tuples, data_provider.py:81) as two device arrays."""
        if not self.neg_examples:
            return None
        if self._neg_pop is None or self._neg_pop[2] is not self.neg_examples:
            pairs = np.asarray(self.neg_examples, dtype=np.int64).reshape(-1, 2)
            if pairs[:, 0].max() >= self._num_users or pairs[:, 1].max() >= self._num_items or pairs.min() < 0:
                raise IndexError('index out of range in self')       # what nn.Embedding raises in the reference
            dev = next(self._net.parameters()).device
            self._neg_pop = (torch.from_numpy(np.ascontiguousarray(pairs[:, 0])).to(dev),
                             torch.from_numpy(np.ascontiguousarray(pairs[:, 1])).to(dev), self.neg_examples)
        return self._neg_pop

    def fit(self, train_set, valid_set, verbose=False):
        self.train_set = train_set
        user_ids = train_set.user_ids
        item_ids = train_set.item_ids
        users, items = shuffle(user_ids, item_ids, random_state=self._random_state)   # once per fit

        if not self._initialized or (self._engine_ is None and not self._generic) or self._optimizer is None:
            self._initialize(train_set)
        self._check_input(user_ids, item_ids)
        if self._generic:
            return self._fit_generic(users, items, valid_set, verbose)
        dev = self._engine_.device
        users_d = torch.from_numpy(np.ascontiguousarray(users)).to(dev).long()
        items_d = torch.from_numpy(np.ascontiguousarray(items)).to(dev).long()
        val_users_d = torch.from_numpy(np.ascontiguousarray(valid_set.user_ids)).to(dev).long()
        val_items_d = torch.from_numpy(np.ascontiguousarray(valid_set.item_ids)).to(dev).long()
        B = self._batch_size
        pop = self._negative_population()
        n_neg = self._num_negative_samples if pop is not None else 0
        pop_u, pop_i = (pop[0], pop[1]) if pop is not None else (None, None)

        total_losses = {"train_loss": [], "validation_loss": [], "curr_epoch": []}
        for epoch_num in range(self._n_iter):
            # Negatives: random.choices(neg_examples, k = n_neg * batch) per step (implicit.py:352,370),
            # drawn on the device from Python's global MT19937 stream -- training steps first, then the
            # validation steps, exactly the order in which the reference consumes it.
            self._net.train()
            if pop is not None:
                self._engine_.rng_seed(random)
            train_losses = self._engine_.train_epoch(self._loss_kind, users_d, items_d, B, n_neg, pop_u, pop_i)
            self._net.eval()
            val_losses = self._engine_.loss_epoch(self._loss_kind, val_users_d, val_items_d, B, n_neg, pop_u, pop_i)
            # one device->host read per epoch replaces three loss.item() syncs per step (implicit.py:294-298)
            train_steps = [float(x) for x in train_losses.cpu().numpy()]
            val_steps = [float(x) for x in val_losses.cpu().numpy()]
            if pop is not None:
                self._engine_.rng_sync(random)       # `random` now sits where the reference would leave it

            train_epoch_loss = sum(train_steps) / len(train_steps)
            if np.isnan(train_epoch_loss) or train_epoch_loss == 0.0:
                raise ValueError('Degenerate epoch loss: {}'.format(train_epoch_loss))
            valid_epoch_loss = sum(val_steps) / len(val_steps)
            if self.best_validation is None or valid_epoch_loss < self.best_validation:
                self._engine_.flush()                       # tables current before the snapshot
                self.best_model = copy.deepcopy(self._net)
                self.best_validation = valid_epoch_loss
                self.best_epoch = epoch_num
            if verbose:
                logging.info('Epoch {}: training_loss {:10.5f}'.format(epoch_num, train_epoch_loss))
                logging.info('Epoch {}: validation_loss {:10.5f}'.format(epoch_num, valid_epoch_loss))
            total_losses["train_loss"].append(np.mean(train_steps))
            total_losses["validation_loss"].append(np.mean(val_steps))
            total_losses['curr_epoch'].append(epoch_num)
            save_statistics(experiment_log_dir=self.experiment_logs, filename='summary.csv',
                            stats_dict=total_losses, current_epoch=epoch_num,
                            continue_from_mode=(self.starting_epoch != 0 or epoch_num > 0))

        self._net = self.best_model
        self._engine_ = None                                # best_model gets a forward-only engine on demand
        self.save_readable_model(self.experiment_saved_models, self.best_model.state_dict())
        logging.info("Model chosen from epoch %d", self.best_epoch)

    # -------------------------------------------------------------------------------------------
    # generic representation (MLP, NeuMF, any torch module): implicit.py:279-343 with torch autograd
    def _generic_step(self, batch_user, batch_item, neg_user, neg_item, train):
        """run_train_iteration / run_val_iteration (implicit.py:347-379) on a torch module.  Module outputs of shape
        [n, 1] (MLP, NeuMF) are flattened: every loss of losses.py treats [n, 1] and [n] alike."""
        pos = self._net(batch_user, batch_item).reshape(-1)
        if train:
            self._optimizer.zero_grad()
        if neg_user is not None:
            neg = self._net(neg_user, neg_item).reshape(-1)
            loss = self._loss_func(pos, neg)
        else:
            loss = self._loss_func(pos)
        if train:
            loss.backward()
            self._optimizer.step()
        return loss.detach()

    def _fit_generic(self, users, items, valid_set, verbose):
        from recommendation_gans_b200.engine import draw_negative_pairs_device
        dev = next(self._net.parameters()).device
        users_d = torch.from_numpy(np.ascontiguousarray(users)).to(dev).long()
        items_d = torch.from_numpy(np.ascontiguousarray(items)).to(dev).long()
        val_users_d = torch.from_numpy(np.ascontiguousarray(valid_set.user_ids)).to(dev).long()
        val_items_d = torch.from_numpy(np.ascontiguousarray(valid_set.item_ids)).to(dev).long()
        B = self._batch_size
        pop = self._negative_population()
        k = self._num_negative_samples * B if pop is not None else 0
        total_losses = {"train_loss": [], "validation_loss": [], "curr_epoch": []}

        def run(ids_u, ids_i, train):
            # all negative pairs of the pass in one device draw: the same stream positions as one
            # random.choices(neg_examples, k) per minibatch (implicit.py:352,370)
            nsteps = (len(ids_u) + B - 1) // B
            neg_u = neg_i = None
            if pop is not None:
                neg_u, neg_i = draw_negative_pairs_device(pop[0], pop[1], nsteps * k, rng=random)
            losses = []
            for s in range(nsteps):
                nu = neg_u[s * k:(s + 1) * k] if pop is not None else None
                ni = neg_i[s * k:(s + 1) * k] if pop is not None else None
                losses.append(self._generic_step(ids_u[s * B:(s + 1) * B], ids_i[s * B:(s + 1) * B], nu, ni, train))
            return [float(x) for x in torch.stack(losses).cpu().numpy()]     # one device->host read per pass

        for epoch_num in range(self._n_iter):
            self._net.train()
            train_steps = run(users_d, items_d, True)
            train_epoch_loss = sum(train_steps) / len(train_steps)
            if np.isnan(train_epoch_loss) or train_epoch_loss == 0.0:
                raise ValueError('Degenerate epoch loss: {}'.format(train_epoch_loss))
            self._net.eval()
            with torch.no_grad():
                val_steps = run(val_users_d, val_items_d, False)
            valid_epoch_loss = sum(val_steps) / len(val_steps)
            if self.best_validation is None or valid_epoch_loss < self.best_validation:
                self.best_model = copy.deepcopy(self._net)
                self.best_validation = valid_epoch_loss
                self.best_epoch = epoch_num
            if verbose:
                logging.info('Epoch {}: training_loss {:10.5f}'.format(epoch_num, train_epoch_loss))
                logging.info('Epoch {}: validation_loss {:10.5f}'.format(epoch_num, valid_epoch_loss))
            total_losses["train_loss"].append(np.mean(train_steps))
            total_losses["validation_loss"].append(np.mean(val_steps))
            total_losses['curr_epoch'].append(epoch_num)
            save_statistics(experiment_log_dir=self.experiment_logs, filename='summary.csv',
                            stats_dict=total_losses, current_epoch=epoch_num,
                            continue_from_mode=(self.starting_epoch != 0 or epoch_num > 0))
        self._net = self.best_model
        self._optimizer = None                              # bound to the pre-copy parameters: rebuilt by the next fit
        self.save_readable_model(self.experiment_saved_models, self.best_model.state_dict())
        logging.info("Model chosen from epoch %d", self.best_epoch)

    # single-step API (implicit.py:347-379): same kernels, one minibatch
    def _one_batch(self, batch_user, batch_item, train):
        if self._generic:
            from recommendation_gans_b200.engine import draw_negative_pairs_device
            if self._optimizer is None:
                raise RuntimeError('model is not initialised for training (call fit or _initialize first)')
            dev = next(self._net.parameters()).device
            pop = self._negative_population()
            nu = ni = None
            if pop is not None:
                nu, ni = draw_negative_pairs_device(pop[0], pop[1], self._num_negative_samples * self._batch_size,
                                                    rng=random)
            bu = torch.as_tensor(batch_user, device=dev).long()
            bi = torch.as_tensor(batch_item, device=dev).long()
            if train:
                return self._generic_step(bu, bi, nu, ni, True)
            with torch.no_grad():
                return self._generic_step(bu, bi, nu, ni, False)
        if self._engine_ is None:
            raise RuntimeError('model is not initialised for training (call fit or _initialize first)')
        if len(batch_user) > self._batch_size:
            raise ValueError('batch larger than batch_size')
        pop = self._negative_population()
        n_neg, neg_u, neg_i = 0, None, None
        if pop is not None:
            n_neg = self._num_negative_samples
            neg_u, neg_i = self._engine_.draw_negative_pairs(pop[0], pop[1], n_neg * self._batch_size, rng=random)
        # n_pos = len(batch) < batch_size makes this the (partial) final step of a one-step epoch:
        # b positives against n_neg * batch_size negatives, exactly as implicit.py:352 sizes k.
        # hinge on a partial batch raises RuntimeError like torch's broadcast in losses.py:121.
        fn = self._engine_.train_steps if train else self._engine_.loss_steps
        return fn(self._loss_kind, batch_user, batch_item, self._batch_size, n_neg, neg_u, neg_i)[0]

    def run_train_iteration(self, batch_user, batch_item):
        loss = self._one_batch(batch_user, batch_item, True)
        if self._engine_ is not None:
            self._engine_.flush()       # the reference leaves net.*.weight current after every iteration
        return loss

    def run_val_iteration(self, batch_user, batch_item):
        return self._one_batch(batch_user, batch_item, False)

    # -------------------------------------------------------------------------------------------
    def predict(self, user_ids, item_ids=None):
        self._check_input(user_ids, item_ids, allow_items_none=True)
        self._net.train(False)
        if self._generic:
            users, items = _predict_process_ids(user_ids, item_ids, self._num_items, True)
            with torch.no_grad():
                out = self._net(users, items)
            return out.cpu().detach().numpy().flatten()
        eng = self._engine_ if self._engine_ is not None else self._net._engine()
        if np.isscalar(user_ids) and item_ids is None:
            out = eng.predict_user(int(user_ids))
        else:
            users, items = _predict_process_ids(user_ids, item_ids, self._num_items, True)
            out = eng.predict_pairs(users, items)
        return out.cpu().numpy().flatten()

    def test(self, test_set, item_popularity, k=5, rmse_flag=False, precision_recall=False, map_recall=True):
        from spotlight.evaluation import evaluate_popItems, evaluate_random, map_at_k
        self._net.eval()
        dev = next(self._net.parameters()).device
        test_users = torch.from_numpy(np.ascontiguousarray(test_set.user_ids)).to(dev).long()
        test_items = torch.from_numpy(np.ascontiguousarray(test_set.item_ids)).to(dev).long()
        test_results = {'k': k}
        if rmse_flag:
            total = 0
            for batch_user, batch_item in minibatch(test_users, test_items, batch_size=self._batch_size):
                total += rmse_score(self._net, batch_user, batch_item)
            total /= len(test_set)
            logging.info("BCE: {}".format(np.sqrt(total)))
            test_results["bce"] = float(np.sqrt(total))     # float(): json cannot serialise np.float32 (F10)
        if precision_recall:
            pop_precision, pop_recall = evaluate_popItems(item_popularity, test_set, k=k)
            rand_precision, rand_recall = evaluate_random(item_popularity, test_set, k=k)
            precision, recall = precision_recall_score(self, test=test_set, k=k)
            logging.info(self.model_name + " precision@{} {} recall@{} {}".format(k, precision, k, recall))
            logging.info("Random: precision@{} {} recall@{} {}".format(k, rand_precision, k, rand_recall))
            logging.info("PopItem Algorithm: precision@{} {} recall@{} {}".format(k, pop_precision, k, pop_recall))
            test_results.update(precision=precision, recall=recall, rand_prec=rand_precision,
                                rand_rec=rand_recall, pop_prec=pop_precision, pop_rec=pop_recall, at_k=k)
        if map_recall:
            map_k = map_at_k(self, test=test_set, k=k)
            _, recall = precision_recall_score(self, test=test_set, k=k)
            logging.info(self.model_name + " map@{} {} recall@{} {}".format(k, map_k, k, recall))
            test_results["map"] = map_k
        with open(os.path.join(self.experiment_logs, 'test_summary.json'), 'w') as fp:
            json.dump(test_results, fp)
        return test_results

    def save_readable_model(self, model_save_dir, state_dict):
        fname = os.path.join(model_save_dir, "best_model")
        logging.info('Saving state in {}'.format(fname))
        torch.save({'network': state_dict}, f=fname)
