// phycpp_harness.cpp -- TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED C++ wrapper of the reference (src/phycpp/physher.cpp,
// compiled from where it lies) through its own public classes, the way torchtree-physher does: tree model, substitution model,
// site model, TreeLikelihoodInterface, LogLikelihood(), RequestGradient(), Gradient().  Built twice by oracle/Makefile:
//   _ref/libphycpp_cpu.so    plain link against the compiled reference           (the answers)
//   _ref/libphycpp_b200.so   the same objects linked with --wrap (integration/phycpp_wrap.c): the likelihood runs on the device
// tests/test_glue_dropin.py loads both and compares.
#include <cstring>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "phycpp/physher.hpp"

extern "C" {

// unrooted tree, GTR (rates simplex of 6) + Gamma(ncat) or constant rates: returns the gradient length (N - 2), lnL in *lnl,
// d lnL / d branch length in grad[] indexed as TreeLikelihoodInterface::Gradient indexes it (nodeMap_, physher.cpp:644-660)
int phycpp_unrooted_gtr(const char *newick, int ntaxa, const char **taxa, const char **seqs, const double *rates6, const double *freqs4,
                        double alpha, int ncat, int use_tip_states, double *lnl, double *grad, int cap) {
	std::vector<std::string> names(taxa, taxa + ntaxa);
	std::vector<std::pair<std::string, std::string>> aln;
	for (int i = 0; i < ntaxa; i++) aln.emplace_back(taxa[i], seqs[i]);
	UnRootedTreeModelInterface tree(newick, names);
	GTRInterface gtr(std::vector<double>(rates6, rates6 + 6), std::vector<double>(freqs4, freqs4 + 4));
	SiteModelInterface *site = ncat > 1 ? static_cast<SiteModelInterface *>(new GammaSiteModelInterface(alpha, ncat, std::nullopt, std::nullopt))
	                                   : static_cast<SiteModelInterface *>(new ConstantSiteModelInterface(std::nullopt));
	int n = 0;
	{
		TreeLikelihoodInterface tlk(aln, &tree, &gtr, site, std::nullopt, false, use_tip_states != 0, false);
		// explicit flags: the constructor's RequestGradient() passes 0, which under-sizes the gradient buffer for models with dPdp
		tlk.RequestGradient({TreeLikelihoodGradientFlags::TREE_HEIGHT});
		*lnl = tlk.LogLikelihood();
		n = (int)tlk.gradientLength_;
		std::vector<double> g(n + 2, 0.0);
		tlk.Gradient(g.data());
		for (int i = 0; i < n && i < cap; i++) grad[i] = g[i];
	}
	delete site;
	return n;
}

// rooted time tree (tip dates), JC69, strict clock: lnL and the gradient TreeLikelihoodInterface::Gradient returns for
// {tree, branch model} (reparameterised ratios / root height, then the clock rate)
int phycpp_time_jc69(const char *newick, int ntaxa, const char **taxa, const double *dates, const char **seqs, double rate, double *lnl,
                     double *grad, int cap) {
	std::vector<std::string> names(taxa, taxa + ntaxa);
	std::vector<double> d(dates, dates + ntaxa);
	std::vector<std::pair<std::string, std::string>> aln;
	for (int i = 0; i < ntaxa; i++) aln.emplace_back(taxa[i], seqs[i]);
	ReparameterizedTimeTreeModelInterface tree(newick, names, d, TreeTransformFlags::RATIO);
	JC69Interface jc;
	ConstantSiteModelInterface site(std::nullopt);
	StrictClockModelInterface clock(rate, &tree);
	TreeLikelihoodInterface tlk(aln, &tree, &jc, &site, &clock, false, true, false);
	tlk.RequestGradient({TreeLikelihoodGradientFlags::TREE_HEIGHT, TreeLikelihoodGradientFlags::BRANCH_MODEL});
	*lnl = tlk.LogLikelihood();
	const int n = (int)tlk.gradientLength_;
	std::vector<double> g(n + 2, 0.0);
	tlk.Gradient(g.data());
	for (int i = 0; i < n && i < cap; i++) grad[i] = g[i];
	return n;
}
}
