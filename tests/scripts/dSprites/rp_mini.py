"""A flat training script shaped like the reference's dSprites/rp.py (stage 2): a FROZEN but grad-tracked alignment
encoder in eval mode, ``transformation_2D`` built on F.affine_grid + F.grid_sample whose BACKWARD is on the executed
path (through ``torch.inverse(get_matrix_pxy_align(...))`` into the frozen encoder), ``from utils_rp import *`` /
``from utils_pxy import *`` helper modules, a Linear -> view -> ConvTranspose/BatchNorm generator, spectral-norm conv
trunks with spectral-norm Linear heads, ``nn.Softmax()`` with implicit dim, mutual-information loss, two Adams.  Run
UNCHANGED by tests/test_run_gpu.py with stock PyTorch and under ``python -m eadgan_b200.run``."""
import argparse
import itertools
import json

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Variable
from torch.nn.utils import spectral_norm

from utils_rp import *     # noqa: F401,F403
from utils_pxy import *    # noqa: F401,F403

parser = argparse.ArgumentParser()
parser.add_argument("--n_iter", type=int, default=3)
parser.add_argument("--batch_size", type=int, default=16)
parser.add_argument("--code_dim", type=int, default=4)
parser.add_argument("--n_classes", type=int, default=3)
parser.add_argument("--seed", type=int, default=0)
opt = parser.parse_args()
torch.manual_seed(opt.seed)
np.random.seed(opt.seed)
FloatTensor = torch.cuda.FloatTensor


def trunk(sn, slope):
    wrap = spectral_norm if sn else (lambda m: m)
    layers = []
    for cin, cout in ((1, 32), (32, 32), (32, 64), (64, 64)):
        layers += [wrap(nn.Conv2d(cin, cout, 4, 2, 1)), nn.LeakyReLU(slope, inplace=True)]
    return nn.Sequential(*layers)


class Encoder_pxy(nn.Module):
    def __init__(self):
        super(Encoder_pxy, self).__init__()
        self.conv_block = trunk(False, 0.1)
        self.fc1 = nn.Linear(1024, 3)

    def forward(self, img):
        x = self.conv_block(img)
        return self.fc1(x.view(x.shape[0], -1))


class Discriminator(nn.Module):
    def __init__(self):
        super(Discriminator, self).__init__()
        self.conv_block = trunk(True, 0.2)
        self.fc1 = nn.Sequential(spectral_norm(nn.Linear(1024, 128)), nn.LeakyReLU(0.2, inplace=True))
        self.fc2 = nn.Linear(128, 1)

    def forward(self, img):
        x = self.conv_block(img)
        return F.sigmoid(self.fc2(self.fc1(x.view(x.shape[0], -1))))


class Generator(nn.Module):
    def __init__(self):
        super(Generator, self).__init__()
        blocks = []
        for _ in range(3):
            blocks += [nn.ConvTranspose2d(64, 64, 4, 2, 1), nn.BatchNorm2d(64), nn.ReLU()]
        self.conv_block = nn.Sequential(*blocks, nn.ConvTranspose2d(64, 1, 4, 2, 1))
        self.fc1 = nn.Sequential(nn.Linear(opt.n_classes + opt.code_dim, 128), nn.ReLU())
        self.fc2 = nn.Sequential(nn.Linear(128, 64 * 4 * 4), nn.ReLU())

    def forward(self, z_c):
        x = self.fc2(self.fc1(z_c))
        return F.sigmoid(self.conv_block(x.view(x.shape[0], 64, 4, 4)))


class Encoder(nn.Module):
    def __init__(self):
        super(Encoder, self).__init__()
        self.conv_block = trunk(True, 0.2)
        self.fc1 = nn.Sequential(spectral_norm(nn.Linear(1024, 128)), nn.LeakyReLU(0.2, inplace=True))
        self.fc2 = nn.Sequential(spectral_norm(nn.Linear(128, 128)), nn.LeakyReLU(0.2, inplace=True))
        self.cat_layer = nn.Sequential(spectral_norm(nn.Linear(128, opt.n_classes)), nn.Softmax())
        self.cont_layer = nn.Sequential(spectral_norm(nn.Linear(128, opt.code_dim)))

    def forward(self, img):
        x = self.conv_block(img)
        x = self.fc2(self.fc1(x.view(x.shape[0], -1)))
        return self.cat_layer(x), self.cont_layer(x)


class transformation_2D(nn.Module):
    def forward(self, img, matrix_2D):
        grid = F.affine_grid(matrix_2D, img.size())
        return F.grid_sample(img, grid, padding_mode="border")


def mutual_info_loss(c_given_x, c):
    eps = 1e-8
    return torch.mean(-torch.sum(torch.log(c_given_x + eps) * c, dim=1)) + torch.mean(-torch.sum(torch.log(c + eps) * c, dim=1))


def to_categorical(y, num_columns):
    y_cat = np.zeros((y.shape[0], num_columns))
    y_cat[range(y.shape[0]), y] = 1.0
    return Variable(FloatTensor(y_cat))


continuous_loss = torch.nn.MSELoss().cuda()
adv_loss = torch.nn.BCELoss().cuda()
encoder_pxy, encoder, discriminator, generator = Encoder_pxy(), Encoder(), Discriminator(), Generator()
trans_2D = transformation_2D()
for m in (encoder_pxy, encoder, discriminator, generator):
    m.cuda()
encoder_pxy.eval()          # frozen by never being stepped; its parameters stay grad-tracked, as in the reference
optimizer_D = torch.optim.Adam(discriminator.parameters(), lr=0.0002, betas=(0.5, 0.999))
optimizer_info = torch.optim.Adam(itertools.chain(generator.parameters(), encoder.parameters()), lr=0.0001, betas=(0.5, 0.999))

yy, xx = np.mgrid[0:64, 0:64]
for it in range(opt.n_iter):
    B = opt.batch_size
    cx, cy, rad = np.random.uniform(24, 40, B), np.random.uniform(24, 40, B), np.random.uniform(6, 11, B)
    img = torch.from_numpy(((np.abs(xx[None] - cx[:, None, None]) <= rad[:, None, None]) &
                            (np.abs(yy[None] - cy[:, None, None]) <= rad[:, None, None])).astype(np.uint8))
    img = img.unsqueeze(1).cuda().float()
    valid = Variable(FloatTensor(B, 1).fill_(1.0), requires_grad=False)
    fake = Variable(FloatTensor(B, 1).fill_(0.0), requires_grad=False)

    def aligned():
        align_code = encoder_pxy(img)
        inv = torch.inverse(get_matrix_pxy_align(align_code))
        return trans_2D(img, inv[:, 0:2])

    # ---- discriminator phase
    align_img = aligned()
    code_input = Variable(FloatTensor(np.random.uniform(-1, 1, (B, opt.code_dim))))
    label_input = to_categorical(np.random.randint(0, opt.n_classes, B), opt.n_classes)
    trans_img = trans_2D(align_img, get_matrix_D(code_input[:, :4])[:, 0:2])
    gen_img = generator(torch.cat((label_input, code_input), dim=1))
    d_loss = (adv_loss(discriminator(gen_img.detach()), fake) + adv_loss(discriminator(trans_img), valid)) / 2
    optimizer_D.zero_grad()
    d_loss.backward()
    pxy_grad = float(sum(p.grad.abs().sum() for p in encoder_pxy.parameters() if p.grad is not None))
    optimizer_D.step()

    # ---- generator / encoder phase
    code_input = Variable(FloatTensor(np.random.uniform(-1, 1, (B, opt.code_dim))))
    label_input = to_categorical(np.random.randint(0, opt.n_classes, B), opt.n_classes)
    gen_img = generator(torch.cat((label_input, code_input), dim=1))
    rec_cat, rec_cont = encoder(gen_img)
    g_loss = adv_loss(discriminator(gen_img), valid)
    info_loss = mutual_info_loss(rec_cat, label_input) + continuous_loss(rec_cont, code_input)
    align_img = aligned()
    trans_img = trans_2D(align_img, get_matrix(code_input[:, :4])[:, 0:2])
    align_cat, align_cont = encoder(align_img)
    trans_cat, trans_cont = encoder(trans_img)
    affine_loss = continuous_loss(affine_regularzier(align_cont, trans_cont), code_input)
    relative_cat_loss = mutual_info_loss(trans_cat, Variable(align_cat, requires_grad=False))
    total = info_loss + affine_loss + g_loss + relative_cat_loss
    optimizer_info.zero_grad()
    total.backward()
    optimizer_info.step()

    print(json.dumps({"iter": it, "d_loss": d_loss.item(), "g_loss": g_loss.item(), "info_loss": info_loss.item(),
                      "affine_loss": affine_loss.item(), "total": total.item(), "pxy_grad_l1": pxy_grad,
                      "G": type(generator.conv_block).__module__, "opt": type(optimizer_D).__module__,
                      "utils": get_matrix.__module__}), flush=True)
