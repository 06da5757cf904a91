/*
 * phycpp_wrap.c -- puts physher's C++ wrapper (src/phycpp/physher.cpp, the entry torchtree-physher uses) on the device path WITHOUT
 * touching its source: libphycpp is linked with
 *
 *     -Wl,--wrap=new_TreeLikelihoodModel -Wl,--wrap=TreeLikelihood_gradient      (+ this file, + libphysher_glue)
 *
 * so that TreeLikelihoodInterface's constructor (physher.cpp:560-592, 594-629), which builds its model with
 * new_TreeLikelihoodModel, gets a model with the device backend attached, and TreeLikelihoodInterface::Gradient
 * (physher.cpp:644-660), which calls TreeLikelihood_gradient directly instead of going through a function pointer, lands in
 * phb_physher_gradient.  LogLikelihood() is model->logP -> tlk->calculate, the slot the attach re-points.  Everything else in
 * phycpp -- RequestGradient, the nodeMap_ re-indexing, the model interfaces -- runs unchanged.
 *
 * The device ordinal comes from the environment at construction time (PHYSHER_B200_DEVICE, default 0): the wrapped
 * constructor has no argument to carry it.
 */
#include <stdio.h>
#include <stdlib.h>

#include "phyc/treelikelihood.h"

#include "physher_b200.h"

int phb_physher_attach(Model *model, int device);
double *phb_physher_gradient(Model *self);

Model *__real_new_TreeLikelihoodModel(const char *name, SingleTreeLikelihood *tlk, Model *tree, Model *m, Model *sm, Model *bm);

Model *__wrap_new_TreeLikelihoodModel(const char *name, SingleTreeLikelihood *tlk, Model *tree, Model *m, Model *sm, Model *bm) {
	Model *model = __real_new_TreeLikelihoodModel(name, tlk, tree, m, sm, bm);
	const char *dev = getenv("PHYSHER_B200_DEVICE");
	if (phb_physher_attach(model, dev ? atoi(dev) : 0) != 0) {
		fprintf(stderr, "physher_b200: the device path is not available: %s\n", phb_last_error());
		exit(1); /* the reference's error convention; no silent CPU fallback */
	}
	return model;
}

double *__wrap_TreeLikelihood_gradient(Model *model) { return phb_physher_gradient(model); }
