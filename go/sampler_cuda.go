//go:build cuda

// Package sampler — cgo shim that routes grample's Gibbs hot path to libgrample_b200.so.
//
// Drop this file into the reference's `sampler/` directory and build with `-tags cuda`
// (see INTEGRATION.md).  It keeps the exported API that cmd/root.go and cmd/collapse.go drive:
//
//	NewGibbsSimple, NewGibbsCollapsed, (*GibbsCollapsed).Collapse/BlanketSize/FunctionCount,
//	NeighborVarMax, NewChain, (*Chain).AdvanceChain, Chain.TotalSampleCount/LastSample,
//	MergeChains, ChainConvergence, NewConvergenceSampler, NewIdentitySampler, Adapt.
//
// The pure-Go files it replaces (gibbs-simple.go, gibbs-collapsed.go, chain.go, adaptive.go) get
// the build constraint `//go:build !cuda`; sampler.go (interfaces, UniformSampler) stays.
//
// NOTE: this image has no Go toolchain, so this file has been written against
// include/grample_b200.h but never compiled; the C++ mirror of the same layer
// (grample_b200/host/grample.hpp) is what the tests exercise.
package sampler

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../grample_b200 -lgrample_b200
#include <stdlib.h>
#include "grample_b200.h"
*/
import "C"

import (
	"reflect"
	"runtime"
	"sync"

	"github.com/CraigKelly/grample/model"
	"github.com/CraigKelly/grample/rand"
	"github.com/pkg/errors"
)

// Device-side tuning that has no counterpart in the reference.
var (
	// Replicas is the number of device chains behind ONE *Chain (the reference runs one).
	Replicas = 1024
	// Precision selects the sweep arithmetic (C.GB_F64 follows the reference literally).
	Precision = C.GB_F32
	// Device is the CUDA device the process drives (one process per GPU).
	Device = 0
	// RaoBlackwell makes recorded updates add the sampled conditional to every bin of the variable instead of
	// Marginal[value] += 1 (chain.go:235); lower variance, NOT the reference's estimator (GB_F32 / GB_F64 only).
	RaoBlackwell = false
)

func chainFlags() C.uint32_t {
	f := C.uint32_t(C.GB_CHAINS_HISTORY)
	if RaoBlackwell {
		f |= C.GB_CHAINS_RAO_BLACKWELL
	}
	return f
}

func lastErr(what string) error { return errors.Errorf("%s: %s", what, C.GoString(C.gb_last_error())) }

// devModel owns a gb_model handle.
type devModel struct{ h *C.gb_model }

func (d *devModel) free() {
	if d.h != nil {
		C.gb_model_destroy(d.h)
		d.h = nil
	}
}

// flatten turns *model.Model into the CSR arrays of gb_model_create (raw, non-log tables:
// the library applies Function.UseLogSpace's eps rule itself).
func flatten(m *model.Model) (*devModel, error) {
	n := len(m.Vars)
	card := make([]C.int32_t, n)
	fixed := make([]C.int32_t, n)
	for i, v := range m.Vars {
		if i != v.ID {
			return nil, errors.Errorf("Invalid ID for var %s: expected %d but was %d", v.Name, v.ID, i)
		}
		card[i] = C.int32_t(v.Card)
		fixed[i] = C.int32_t(v.FixedVal)
	}
	scopeOff := make([]C.int32_t, 1, len(m.Funcs)+1)
	tabOff := make([]C.int64_t, 1, len(m.Funcs)+1)
	var scope []C.int32_t
	var tables []C.double
	for _, f := range m.Funcs {
		if f.IsLog {
			return nil, errors.Errorf("Function %v is already in log space", f.Name)
		}
		for _, v := range f.Vars {
			scope = append(scope, C.int32_t(v.ID))
		}
		scopeOff = append(scopeOff, C.int32_t(len(scope)))
		for _, t := range f.Table {
			tables = append(tables, C.double(t))
		}
		tabOff = append(tabOff, C.int64_t(len(tables)))
	}
	d := &devModel{}
	rc := C.gb_model_create(C.int32_t(n), &card[0], &fixed[0], C.int32_t(len(m.Funcs)), &scopeOff[0], &scope[0],
		&tabOff[0], &tables[0], C.int(Device), &d.h)
	if rc != 0 {
		return nil, lastErr("Could not flatten model")
	}
	runtime.SetFinalizer(d, (*devModel).free)
	return d, nil
}

// pool is the device population shared by every chain of the process.
type pool struct {
	h  *C.gb_chains
	mu sync.Mutex
}

var thePool = &pool{}

// GibbsSimple keeps the reference's name; the sampling itself happens in the sweep kernels.
type GibbsSimple struct {
	gen *rand.Generator
	pgm *model.Model
	dev *devModel
}

// NewGibbsSimple creates a new sampler (sampler/gibbs-simple.go:25-115).
func NewGibbsSimple(gen *rand.Generator, m *model.Model) (*GibbsSimple, error) {
	if m == nil {
		return nil, errors.New("No model supplied")
	}
	dev, err := flatten(m)
	if err != nil {
		return nil, err
	}
	for _, v := range m.Vars {
		v.State["Selections"] = 0.0
	}
	return &GibbsSimple{gen: gen, pgm: m, dev: dev}, nil
}

// Sample is kept for interface compatibility (FullSampler); single steps are not exposed by the
// device path — chains advance in whole sweeps through AdvanceChain.
func (g *GibbsSimple) Sample(s []int) (int, error) {
	return -1, errors.New("single-step Sample is not available on the CUDA path; use Chain.AdvanceChain")
}

// GibbsCollapsed supports collapsing specified variables (sampler/gibbs-collapsed.go:17-20).
type GibbsCollapsed struct {
	baseSampler *GibbsSimple
	collapses   uint64
}

// NeighborVarMax mirrors sampler/gibbs-collapsed.go:93.
const NeighborVarMax = C.GB_NEIGHBOR_VAR_MAX

// NewGibbsCollapsed creates a new sampler (sampler/gibbs-collapsed.go:23-40).
func NewGibbsCollapsed(gen *rand.Generator, m *model.Model) (*GibbsCollapsed, error) {
	base, err := NewGibbsSimple(gen, m)
	if base == nil {
		return nil, errors.Wrap(err, "Base simple Gibbs sampler could not be created")
	}
	return &GibbsCollapsed{baseSampler: base}, nil
}

// BlanketSize returns the variable's neighborhood size (gibbs-collapsed.go:81-83).
func (g *GibbsCollapsed) BlanketSize(v *model.Variable) int {
	var n C.int32_t
	C.gb_model_blanket_size(g.baseSampler.dev.h, C.int32_t(v.ID), &n)
	return int(n)
}

// FunctionCount returns the variable's factor count (gibbs-collapsed.go:86-88).
func (g *GibbsCollapsed) FunctionCount(v *model.Variable) int {
	var n C.int32_t
	C.gb_model_function_count(g.baseSampler.dev.h, C.int32_t(v.ID), &n)
	return int(n)
}

// Collapse integrates out the variable given by index; < 0 picks one at random
// (gibbs-collapsed.go:98-314).  The new factor set replaces the sampler's device model.
func (g *GibbsCollapsed) Collapse(varIdx int) (*model.Variable, error) {
	base := g.baseSampler
	var v C.int32_t
	marg := make([]C.double, C.GB_MAX_CARD)
	var nm *C.gb_model
	g.collapses++
	seed := C.uint64_t(uint64(base.gen.Int63()))
	if C.gb_model_collapse(base.dev.h, C.int32_t(varIdx), seed, &v, &marg[0], &nm) != 0 {
		return nil, lastErr("Collapse")
	}
	base.dev = &devModel{h: nm}
	runtime.SetFinalizer(base.dev, (*devModel).free)
	dest := base.pgm.Vars[int(v)]
	dest.Collapsed = true
	for k := range dest.Marginal {
		dest.Marginal[k] = float64(marg[k])
	}
	return dest, nil
}

// Sample — see GibbsSimple.Sample.
func (g *GibbsCollapsed) Sample(s []int) (int, error) { return g.baseSampler.Sample(s) }

func devOf(s FullSampler) (*GibbsSimple, error) {
	switch t := s.(type) {
	case *GibbsSimple:
		return t, nil
	case *GibbsCollapsed:
		return t.baseSampler, nil
	}
	return nil, errors.New("the CUDA path needs a GibbsSimple or GibbsCollapsed sampler")
}

// Chain provides functionality around a Gibbs sampler (sampler/chain.go:13-20).  ChainHistory is
// kept on the device as per-chain half-window histograms.
type Chain struct {
	Target            *model.Model
	Sampler           FullSampler
	ConvergenceWindow int
	TotalSampleCount  int64
	LastSample        []int

	group    C.int32_t
	replicas int
}

var nextChainID uint64

// NewChain returns a chain ready to go; it performs burn-in (sampler/chain.go:151-175).
func NewChain(mod *model.Model, samp FullSampler, cw int, burnIn int64) (*Chain, error) {
	base, err := devOf(samp)
	if err != nil {
		return nil, err
	}
	thePool.mu.Lock()
	defer thePool.mu.Unlock()
	first := C.uint64_t(nextChainID)
	nextChainID += uint64((Replicas + 7) / 8 * 8)
	if thePool.h == nil {
		models := []*C.gb_model{base.dev.h}
		counts := []C.int32_t{C.int32_t(Replicas)}
		seed := C.uint64_t(uint64(base.gen.Int63()))
		if C.gb_chains_create(1, &models[0], &counts[0], seed, first, C.int(Precision), chainFlags(),
			C.int(Device), &thePool.h) != 0 {
			return nil, lastErr("Could not create initial chain")
		}
	} else if C.gb_chains_add_group(thePool.h, base.dev.h, C.int32_t(Replicas), first) != 0 {
		return nil, lastErr("Could not create chain")
	}
	var ng, nOrder C.int32_t
	C.gb_chains_n_groups(thePool.h, &ng)
	C.gb_model_schedule(base.dev.h, &nOrder, nil, nil, nil)
	ch := &Chain{Target: mod, Sampler: samp, ConvergenceWindow: cw, LastSample: make([]int, len(mod.Vars)),
		group: ng - 1, replicas: Replicas}
	sweeps := (burnIn + int64(nOrder) - 1) / int64(nOrder) // burnIn counts single-variable steps
	if C.gb_chains_group_sweep(thePool.h, ch.group, C.int64_t(sweeps), 0) != 0 {
		return nil, errors.Wrap(lastErr("sweep"), "Failure during chain burn in")
	}
	return ch, nil
}

// AdvanceChain asynchronously generates one round of samples (sampler/chain.go:180-218): every
// free variable gains ConvergenceWindow+1 recorded samples per device chain.  The goroutine
// locks its OS thread because the library selects the CUDA device per call.
func (c *Chain) AdvanceChain(wg *sync.WaitGroup) error {
	wg.Add(1)
	go func() {
		defer wg.Done()
		runtime.LockOSThread()
		defer runtime.UnlockOSThread()
		thePool.mu.Lock()
		rc := C.gb_chains_group_advance(thePool.h, c.group, C.int32_t(c.ConvergenceWindow))
		thePool.mu.Unlock()
		if rc != 0 || C.gb_chains_synchronize(thePool.h) != 0 {
			panic("Async sample generation failed - cannot continue")
		}
		c.refresh()
	}()
	return nil
}

// refresh copies counts back into Target.Vars[i].Marginal so MergeChains / sol.Error in
// cmd/root.go work unchanged.
func (c *Chain) refresh() {
	var tc C.int32_t
	base, _ := devOf(c.Sampler)
	C.gb_model_total_card(base.dev.h, &tc)
	counts := make([]C.uint64_t, int(tc))
	thePool.mu.Lock()
	C.gb_chains_group_counts(thePool.h, c.group, &counts[0])
	var total C.int64_t
	C.gb_chains_group_info(thePool.h, c.group, nil, &total, nil)
	state := make([]C.int32_t, c.replicas*len(c.Target.Vars))
	C.gb_chains_get_state(thePool.h, c.group, &state[0])
	thePool.mu.Unlock()
	c.TotalSampleCount = int64(total)
	o := 0
	for i, v := range c.Target.Vars {
		if !v.Collapsed {
			for k := range v.Marginal {
				v.Marginal[k] = float64(c.replicas)/float64(v.Card) + float64(counts[o+k])
			}
		}
		o += v.Card
		c.LastSample[i] = int(state[i])
	}
}

// MergeChains is unchanged from the reference (sampler/chain.go:96-148): it only reads
// Target.Vars, which refresh() keeps current.  (Body omitted here: keep the reference's.)

// measureID maps the reference's Measure functions (model/error.go) to the ids of the device
// convergence kernel; Go funcs are only comparable through their code pointers.
func measureID(f Measure) C.int {
	p := reflect.ValueOf(f).Pointer()
	switch p {
	case reflect.ValueOf(model.MaxAbsDiff).Pointer():
		return C.GB_MAX_ABS
	case reflect.ValueOf(model.MeanAbsDiff).Pointer():
		return C.GB_MEAN_ABS
	case reflect.ValueOf(model.JSDivergence).Pointer():
		return C.GB_JS
	}
	return C.GB_HELLINGER
}

// ChainConvergence returns the per-variable convergence score (sampler/chain.go:32-92) computed
// by the device kernel over every chain of the pool.
func ChainConvergence(chains []*Chain, distFunc Measure, mergedVars []*model.Variable) ([]float64, error) {
	if len(chains) < 2 {
		return nil, errors.Errorf("Convergence requires at least 2 chains")
	}
	var err error
	if len(mergedVars) < 1 {
		if mergedVars, err = MergeChains(chains); err != nil {
			return nil, err
		}
	}
	var merged []C.double
	for _, v := range mergedVars {
		for _, p := range v.Marginal {
			merged = append(merged, C.double(p))
		}
	}
	out := make([]C.double, len(mergedVars))
	thePool.mu.Lock()
	rc := C.gb_chains_convergence(thePool.h, measureID(distFunc), &merged[0], &out[0])
	thePool.mu.Unlock()
	if rc != 0 {
		return nil, lastErr("ChainConvergence")
	}
	vals := make([]float64, len(out))
	for i, x := range out {
		vals[i] = float64(x)
	}
	return vals, nil
}

// ConvergenceSampler creates new collapsed chains based on convergence metrics
// (sampler/adaptive.go:28-52).
type ConvergenceSampler struct {
	BaseModel *model.Model
	DistFunc  Measure
	Gen       *rand.Generator
	MaxChains int
	base      *devModel
}

// NewConvergenceSampler mirrors sampler/adaptive.go:36-52.
func NewConvergenceSampler(gen *rand.Generator, m *model.Model, d Measure) (*ConvergenceSampler, error) {
	if m == nil {
		return nil, errors.Errorf("A full model is required for Adaptation")
	}
	if d == nil {
		d = model.HellingerDiff
	}
	dev, err := flatten(m)
	if err != nil {
		return nil, err
	}
	return &ConvergenceSampler{BaseModel: m, DistFunc: d, Gen: gen, MaxChains: 128, base: dev}, nil
}

// Adapt creates new chains with collapsed variables (sampler/adaptive.go:57-157) in one call:
// candidate filter, convergence scores, lowest-score selection, collapse and chain creation all
// happen behind gb_chains_adapt.
func (c *ConvergenceSampler) Adapt(chains []*Chain, newChainCount int) ([]*Chain, error) {
	if len(chains) < 2 {
		return nil, errors.Errorf("At least 2 chains required for adaptation")
	}
	if len(chains) >= c.MaxChains {
		return chains, nil
	}
	last := chains[len(chains)-1]
	chosen := make([]C.int32_t, newChainCount+1)
	var n C.int32_t
	thePool.mu.Lock()
	first := C.uint64_t(nextChainID)
	rc := C.gb_chains_adapt(thePool.h, c.base.h, C.int32_t(newChainCount), C.int32_t(last.replicas),
		measureID(c.DistFunc), C.int32_t(last.ConvergenceWindow), C.int32_t(c.MaxChains), first, &chosen[0], &n)
	nextChainID += uint64(n) * uint64((last.replicas+7)/8*8)
	var ng C.int32_t
	C.gb_chains_n_groups(thePool.h, &ng)
	thePool.mu.Unlock()
	if rc != 0 {
		return nil, lastErr("Adapt")
	}
	for i := 0; i < int(n); i++ {
		mod := c.BaseModel.Clone()
		v := mod.Vars[int(chosen[i])]
		v.Collapsed = true
		var gm *C.gb_model
		grp := ng - n + C.int32_t(i)
		C.gb_chains_group_info(thePool.h, grp, nil, nil, &gm)
		samp := &GibbsCollapsed{baseSampler: &GibbsSimple{gen: c.Gen, pgm: mod, dev: &devModel{h: gm}}}
		ch := &Chain{Target: mod, Sampler: samp, ConvergenceWindow: last.ConvergenceWindow,
			LastSample: make([]int, len(mod.Vars)), group: grp, replicas: last.replicas}
		chains = append(chains, ch)
	}
	// the new chains' collapsed variables receive their exact local marginals from the merged view
	var tc C.int32_t
	C.gb_model_total_card(c.base.h, &tc)
	merged := make([]C.double, int(tc))
	thePool.mu.Lock()
	rc = C.gb_chains_merged_marginals(thePool.h, &merged[0], nil)
	thePool.mu.Unlock()
	if rc != 0 {
		return nil, lastErr("MergeChains")
	}
	for _, ch := range chains[len(chains)-int(n):] {
		o := 0
		for _, v := range ch.Target.Vars {
			if v.Collapsed {
				for k := range v.Marginal {
					v.Marginal[k] = float64(merged[o+k])
				}
			}
			o += v.Card
		}
	}
	return chains, nil
}

// IdentitySampler is just a non-adaptive strategy (sampler/adaptive.go:13-24).
type IdentitySampler struct{}

// NewIdentitySampler creates a new IdentitySampler.
func NewIdentitySampler() (*IdentitySampler, error) { return &IdentitySampler{}, nil }

// Adapt is the identity.
func (i *IdentitySampler) Adapt(chains []*Chain, newChainCount int) ([]*Chain, error) { return chains, nil }
