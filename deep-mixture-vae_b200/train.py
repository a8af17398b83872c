"""Training driver with the flags of code/train.py (same names and defaults, train.py:28-97) plus a few new ones
(--batch_size, --gemm_dtype, --seed, --hidden, --data_n; defaults reproduce the reference).  Plotting / visdom are not
part of the accelerated path: the flags are accepted and ignored with a note.

    python -m dmvae_b200.train --model dmvae --dataset synthetic_mnist --n_epochs 3
"""
from __future__ import annotations

import argparse
import math
import os

import numpy as np

parser = argparse.ArgumentParser(description="Training file for DMVAE and DVMOE")

parser.add_argument("--model", type=str, default="dmvae", help="Model to use [dmvae, vade, dmoe, dvmoe, vademoe]")
parser.add_argument("--model_name", type=str, default="", help="Name of the model")
parser.add_argument("--dataset", type=str, default="mnist", help="Dataset to use [mnist, spiral, cifar10]")
parser.add_argument("--latent_dim", type=int, default=10, help="Number of dimensions for latent variable Z")
parser.add_argument("--output_dim", type=int, default=1, help="Output dimension for regression variable for ME models")
parser.add_argument("--n_clusters", type=int, default=-1, help="Number of clusters to use")
parser.add_argument("--n_experts", type=int, default=5, help="Number of experts to use for MoE models")
parser.add_argument("--classification", action="store_true", default=False,
                    help="Whether the objective is classification or regression (ME models)")
parser.add_argument("--n_epochs", type=int, default=500, help="Number of epochs for training the model")
parser.add_argument("--pretrain_epochs_vae", type=int, default=200, help="Number of epochs for pretraining the vae model")
parser.add_argument("--pretrain_epochs_prior", type=int, default=200, help="Number of epochs for pretraining the gmm model")
parser.add_argument("--init_lr", type=float, default=0.002, help="Initial learning rate for training")
parser.add_argument("--decay_rate", type=float, default=0.9, help="Decay rate for exponentially decaying learning rate (< 1.0)")
parser.add_argument("--decay_epochs", type=int, default=25, help="Number of epochs between exponentially decay of learning rate")
parser.add_argument("--pretrain", action="store_true", default=False, help="Whether to pretrain the model or not")
parser.add_argument("--pretrain_vae_lr", type=float, default=0.0005, help="Initial learning rate for pretraining the vae")
parser.add_argument("--pretrain_decay_rate", type=float, default=0.9,
                    help="Decay rate for exponentially decaying learning rate (< 1.0) for pretraining")
parser.add_argument("--pretrain_decay_epochs", type=int, default=25,
                    help="Number of epochs between exponentially decay of learning rate for pretraining")
parser.add_argument("--pretrain_prior_lr", type=float, default=0.0005, help="Initial learning rate for pretraining the prior")
parser.add_argument("--kl_annealing", action="store_true", default=False, help="Whether to anneal the KL term while training or not")
parser.add_argument("--anneal_step", type=float, default=0.1, help="Step size for annealing")
parser.add_argument("--anneal_epochs", type=int, default=1000, help="Number of epochs before annealing the KL term")
parser.add_argument("--plotting", action="store_true", default=False, help="Whether to generate sampling and regeneration plots")
parser.add_argument("--plot_epochs", type=int, default=100, help="Nummber of epochs before generating plots")
parser.add_argument("--save_epochs", type=int, default=10, help="Nummber of epochs before saving model")
parser.add_argument("--debug", action="store_true", default=False, help="Whether to debug the models or not")
parser.add_argument("--visdom", action="store_true", default=False, help="Using visdom for plotting")
parser.add_argument("--featLearn", action="store_true", default=False, help="Whether to use feature learning in MOE")
# ---- new flags (defaults reproduce the reference) ----
parser.add_argument("--batch_size", type=int, default=100, help="Batch size (hard-coded to 100 in the reference, train.py:215-216)")
parser.add_argument("--gemm_dtype", type=str, default="bf16", help="bf16 (tcgen05 tensor cores) or fp32 (exact tier)")
parser.add_argument("--seed", type=int, default=None, help="Seed for NumPy and the parameter initialisation")
parser.add_argument("--data_n", type=int, default=None, help="Number of training rows for the synthetic datasets")


def main(argv):
    from . import base_models, models, nn
    from .includes.utils import Dataset, MEDataset, load_data
    from .session import Session, global_variables_initializer

    if argv.seed is not None:
        np.random.seed(argv.seed)
    model_str, model_name = argv.model, argv.model_name
    moe = model_str[-3:] == "moe"
    extra = {}
    if argv.data_n is not None and (argv.dataset.startswith("synthetic") or argv.dataset in ("mnist", "cifar10")):
        extra = dict(n_train=argv.data_n, n_test=max(1000, argv.data_n // 5))
    dataset = load_data(argv.dataset, classification=argv.classification, output_dim=argv.output_dim, **extra)
    print(dataset.input_type)
    if model_name == "":
        model_name = model_str
    output_dim = argv.output_dim

    if moe:
        if argv.classification:
            output_dim = dataset.n_classes
        if model_str not in ["dmoe", "vademoe", "dvmoe"]:
            raise NotImplementedError
        if model_str == "dmoe":
            model = models.DeepMoE(model_str, dataset.input_type, dataset.input_dim, output_dim, argv.n_experts,
                                   argv.classification, activation=nn.relu, initializer=nn.xavier_initializer,
                                   featLearn=argv.featLearn).build_graph()
        elif model_str == "dvmoe":
            model = models.DeepVariationalMoE(model_str, dataset.input_type, dataset.input_dim, argv.latent_dim, output_dim,
                                              argv.n_experts, argv.classification, activation=nn.relu,
                                              initializer=nn.xavier_initializer, featLearn=argv.featLearn).build_graph()
        else:
            model = models.VaDEMoE(model_str, dataset.input_type, dataset.input_dim, argv.latent_dim, output_dim,
                                   argv.n_experts, argv.classification, activation=nn.relu,
                                   initializer=nn.xavier_initializer, featLearn=argv.featLearn).build_graph()
        test_data = (dataset.test_data, dataset.test_classes, dataset.test_labels)
        train_data = (dataset.train_data, dataset.train_classes, dataset.train_labels)
        DS = MEDataset
    else:
        n_clusters = argv.n_clusters
        if n_clusters < 1:
            n_clusters = dataset.n_classes
        if model_str not in ["dmvae", "vade"]:
            raise NotImplementedError
        cls = base_models.DeepMixtureVAE if model_str == "dmvae" else base_models.VaDE
        model = cls(model_name, dataset.input_type, dataset.input_dim, argv.latent_dim, n_clusters, activation=nn.relu,
                    initializer=nn.xavier_initializer).build_graph()
        train_data = (np.concatenate([dataset.train_data, dataset.test_data], axis=0),
                      np.concatenate([dataset.train_classes, dataset.test_classes], axis=0))      # train.py:205-207
        test_data = (dataset.test_data, dataset.test_classes)
        DS = Dataset
    model.gemm_dtype = argv.gemm_dtype
    if argv.seed is not None:
        (model.vae if moe else model).seed = argv.seed

    test_data = DS(test_data, batch_size=argv.batch_size)
    train_data = DS(train_data, batch_size=argv.batch_size)
    model.define_train_step(argv.init_lr, train_data.epoch_len * argv.decay_epochs, argv.decay_rate)
    if argv.pretrain:
        if model_str in ["dvmoe", "vademoe"]:
            model.define_pretrain_step(argv.pretrain_vae_lr, train_data.epoch_len * argv.pretrain_decay_epochs,
                                       argv.pretrain_decay_rate)
        elif model_str in ["dmvae", "vade"]:
            model.define_pretrain_step(argv.pretrain_vae_lr, argv.pretrain_prior_lr)

    model.path = "saved-models/%s/%s" % (dataset.datagroup, model.name)
    for path in [model.path + "/" + x for x in ["model", "vae", "prior"]]:
        os.makedirs(path, exist_ok=True)

    sess = Session()
    global_variables_initializer().run(session=sess)
    if argv.pretrain:
        if model_str in ["dvmoe", "vademoe"]:
            model.pretrain(sess, train_data, argv.pretrain_epochs_vae)
        elif model_str in ["dmvae", "vade"]:
            model.pretrain(sess, train_data, argv.pretrain_epochs_vae, argv.pretrain_epochs_prior)

    saver = model.vae if moe else model
    ckpt_path = model.path + "/model/parameters.ckpt"
    try:
        saver.restore(ckpt_path)
    except Exception:
        print("Could not load trained model")
    if argv.visdom or argv.plotting:
        print("note: --visdom / --plotting are outside the accelerated path and are ignored")

    from tqdm import tqdm
    maxAcc = 0.0
    with tqdm(range(argv.n_epochs), postfix={"loss": "inf", "accTrain": "0.00%", "accTest": "0.00%"}) as bar:
        anneal_term = 0.0 if argv.kl_annealing else 1.0
        for epoch in bar:
            if argv.kl_annealing and (epoch + 1) % argv.anneal_epochs == 0:
                anneal_term = min(anneal_term + argv.anneal_step, 1.0)
            if moe:
                loss, accTrain, lossCls = model.train_op(sess, train_data, anneal_term)
                accTest, accClsTest = model.get_accuracy(sess, test_data)
            else:
                loss = model.train_op(sess, train_data, anneal_term)
                accTrain = model.get_accuracy(sess, train_data)
                accTest = model.get_accuracy(sess, test_data)
                accClsTest = accTest
            if accTest > maxAcc:
                maxAcc = accTest
                saver.save(ckpt_path)
            if math.isnan(loss):
                raise FloatingPointError("loss is NaN at epoch %d" % epoch)          # the reference drops into pdb here
            bar.set_postfix({"loss": "%.4f" % loss, "accTrain": "%.4f" % accTrain, "accTest": "%.4f" % accTest,
                             "maxAcc": "%.4f" % maxAcc, "accClusteringTest": "%.4f" % accClsTest})
    with open(argv.model + "_logs.txt", "a+") as fl:
        fl.write("\n" + str(argv) + "\n------\n")
        fl.write("Max Accuracy        " + str(maxAcc) + "\n============")
    return maxAcc


if __name__ == "__main__":
    args = parser.parse_args()
    print(args)
    main(args)
