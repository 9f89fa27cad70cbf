//go:build cuda

// Package sampler — cgo shim that routes grample's Gibbs hot path to libgrample_b200.so.
//
// Drop this file into the reference's `sampler/` directory and build with `-tags cuda`
// (see INTEGRATION.md).  It keeps the exported API that cmd/root.go and cmd/collapse.go drive:
//
//	NewGibbsSimple, (*GibbsSimple).Sample/SampleVar/FunctionsChanged, NewGibbsCollapsed,
//	(*GibbsCollapsed).Collapse/BlanketSize/FunctionCount/Sample/FunctionsChanged, NeighborVarMax,
//	Measure, NewChain, (*Chain).AdvanceChain/ChainDist, Chain.TotalSampleCount/LastSample/ChainHistory,
//	MergeChains, ChainConvergence, NewConvergenceSampler, NewIdentitySampler, Adapt.
//
// The pure-Go files it replaces (gibbs-simple.go, gibbs-collapsed.go, chain.go, adaptive.go) get
// the build constraint `//go:build !cuda`; sampler.go (interfaces, UniformSampler) stays.  With
// chain.go excluded this file also defines what lived there and is not sampling: type Measure
// (chain.go:24), MergeChains (chain.go:96-148) and (*Chain).ChainDist (chain.go:253-290).
//
// Concurrency: every C-ABI entry that takes the gb_chains handle holds the handle's own lock
// (include/grample_b200.h, "thread safety"), so the goroutines AdvanceChain spawns may call in
// concurrently; thePool.mu only guards the creation of the handle and the chain-id counter.
//
// NOTE: this image has no Go toolchain, so this file has been written against
// include/grample_b200.h but never compiled; the C++ mirror of the same layer
// (grample_b200/host/grample.hpp) is what the tests exercise, and tests/threads_test.cpp makes
// this file's concurrent call pattern (16 threads on one handle) through the C ABI.
package sampler

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../grample_b200 -lgrample_b200
#include <stdlib.h>
#include "grample_b200.h"
*/
import "C"

import (
	"math"
	"reflect"
	"runtime"
	"sync"
	"sync/atomic"

	"github.com/CraigKelly/grample/buffer"
	"github.com/CraigKelly/grample/model"
	"github.com/CraigKelly/grample/rand"
	"github.com/pkg/errors"
)

// Device-side tuning that has no counterpart in the reference.
var (
	// Replicas is the number of device chains behind ONE *Chain (the reference runs one).
	Replicas = 1024
	// Precision selects the sweep arithmetic.  GB_HYBRID is the one default of every host of the
	// boundary (this shim, the C++ mirror, the CLI's `auto`): the reference's float64 arithmetic
	// throughout, sampled from float64-derived threshold tables where a variable's conditional can
	// be tabulated and by the float64 log-sum-exp kernels elsewhere.  C.GB_F64 forces the
	// per-update log-sum-exp everywhere; C.GB_F32 is the float32 opt-in.
	Precision = C.GB_HYBRID
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

// chainPrecision: the Rao-Blackwell estimator is available under every precision the shim offers
// (tabulated conditionals are read back from their thresholds).
func chainPrecision() C.int { return C.int(Precision) }

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

// GibbsSimple keeps the reference's name; chains advance in the sweep kernels, single steps
// (Sample / SampleVar) go through gb_model_sample.
type GibbsSimple struct {
	gen  *rand.Generator
	pgm  *model.Model
	dev  *devModel
	seed uint64 // Philox key of the single-step path
	step uint64 // its counter (one per Sample / SampleVar call)
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
	return &GibbsSimple{gen: gen, pgm: m, dev: dev, seed: uint64(gen.Int63())}, nil
}

// FunctionsChanged must be called after the model's Funcs changed (gibbs-simple.go:119-145): the
// device copy is rebuilt from the Go model.  (Collapse does not need it: the library returns the
// collapsed model itself.)  The reference also re-draws its private state here; on this path a
// chain's state lives in its device group, which restarts when NewChain is called on the sampler.
func (g *GibbsSimple) FunctionsChanged() error {
	dev, err := flatten(g.pgm)
	if err != nil {
		return err
	}
	g.dev = dev
	return nil
}

func (g *GibbsSimple) sampleOne(varIdx int, excludeCollapsed bool, s []int) (int, error) {
	if len(s) != len(g.pgm.Vars) {
		return -1, errors.Errorf("Sample size %d != Var size %d in model %s", len(s), len(g.pgm.Vars), g.pgm.Name)
	}
	st := make([]C.int32_t, len(s))
	for i, x := range s {
		st[i] = C.int32_t(x)
	}
	excl := C.int(0)
	if excludeCollapsed {
		excl = 1
	}
	prec := C.int(C.GB_F64)
	if Precision == C.GB_F32 {
		prec = C.GB_F32
	}
	var v C.int32_t
	step := atomic.AddUint64(&g.step, 1)
	if C.gb_model_sample(g.dev.h, prec, C.int32_t(varIdx), excl, C.uint64_t(g.seed), C.uint64_t(step), &st[0], &v) != 0 {
		return -1, lastErr("Could not sample from var in model")
	}
	g.pgm.Vars[int(v)].State["Selections"] += 1.0 // gibbs-simple.go:165
	s[int(v)] = int(st[int(v)])
	return int(v), nil
}

// Sample returns a single sample — uniformly selects a variable to sample from
// (gibbs-simple.go:148-160; FullSampler).  One small kernel launch per call: the throughput path is
// Chain.AdvanceChain.
func (g *GibbsSimple) Sample(s []int) (int, error) { return g.sampleOne(-1, false, s) }

// SampleVar samples the given variable conditioned on the rest of `s` (gibbs-simple.go:163-271).
func (g *GibbsSimple) SampleVar(varIdx int, s []int) (int, error) {
	if varIdx < 0 || varIdx >= len(g.pgm.Vars) {
		return -1, errors.Errorf("Invalid variable index %d", varIdx)
	}
	return g.sampleOne(varIdx, false, s)
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

// Sample is GibbsSimple.Sample over the un-collapsed variables (gibbs-collapsed.go:317-334).
func (g *GibbsCollapsed) Sample(s []int) (int, error) { return g.baseSampler.sampleOne(-1, true, s) }

// FunctionsChanged — see GibbsSimple.FunctionsChanged (gibbs-collapsed.go:44-78: the neighbour sets
// are rebuilt by gb_model_create).
func (g *GibbsCollapsed) FunctionsChanged() error { return g.baseSampler.FunctionsChanged() }

func devOf(s FullSampler) (*GibbsSimple, error) {
	switch t := s.(type) {
	case *GibbsSimple:
		return t, nil
	case *GibbsCollapsed:
		return t.baseSampler, nil
	}
	return nil, errors.New("the CUDA path needs a GibbsSimple or GibbsCollapsed sampler")
}

// Measure is an error metric used by ChainConvergence, e.g. model.HellingerDiff (sampler/chain.go:24).
type Measure func(v1 *model.Variable, v2 *model.Variable) float64

// Chain provides functionality around a Gibbs sampler (sampler/chain.go:13-20).  The sliding windows
// live on the device as per-replica half-window histograms; ChainHistory is filled after every round
// with replica 0's window (values in ascending order inside each half — every consumer in the
// reference only counts them) so that code reading it keeps working.
type Chain struct {
	Target            *model.Model
	Sampler           FullSampler
	ConvergenceWindow int
	ChainHistory      []*buffer.CircularInt
	TotalSampleCount  int64
	LastSample        []int

	group    C.int32_t
	replicas int
	hist     []C.uint16_t // [2][sum(card)][replicas] half-window histograms of the last round
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
		ChainHistory: make([]*buffer.CircularInt, len(mod.Vars)), group: ng - 1, replicas: Replicas}
	for i := range ch.ChainHistory {
		ch.ChainHistory[i] = buffer.NewCircularInt(cw)
	}
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
		runtime.LockOSThread() // gb_last_error is thread-local
		defer runtime.UnlockOSThread()
		// no shim-level lock: the handle serialises concurrent callers itself, and
		// gb_chains_synchronize waits outside the handle's lock
		if C.gb_chains_group_advance(thePool.h, c.group, C.int32_t(c.ConvergenceWindow)) != 0 ||
			C.gb_chains_synchronize(thePool.h) != 0 {
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
	C.gb_chains_group_counts(thePool.h, c.group, &counts[0])
	var total C.int64_t
	C.gb_chains_group_info(thePool.h, c.group, nil, &total, nil)
	state := make([]C.int32_t, c.replicas*len(c.Target.Vars))
	C.gb_chains_get_state(thePool.h, c.group, &state[0])
	c.hist = make([]C.uint16_t, 2*int(tc)*c.replicas)
	C.gb_chains_group_history(thePool.h, c.group, &c.hist[0])
	c.TotalSampleCount = int64(total)
	unit := 1.0
	if RaoBlackwell {
		unit = 1.0 / 16777216.0 // bins are fixed point in units of 2^-24
	}
	o := 0
	for i, v := range c.Target.Vars {
		if !v.Collapsed {
			for k := range v.Marginal {
				v.Marginal[k] = float64(c.replicas)/float64(v.Card) + float64(counts[o+k])*unit
			}
		}
		// replica 0's window, oldest half first (buffer/circular.go: Add keeps the newest BufSize values)
		if v.FixedVal < 0 && !v.Collapsed {
			for half := 0; half < 2; half++ {
				for k := 0; k < v.Card; k++ {
					n := int(c.hist[(half*int(tc)+o+k)*c.replicas])
					for j := 0; j < n; j++ {
						c.ChainHistory[i].Add(k)
					}
				}
			}
		}
		o += v.Card
		c.LastSample[i] = int(state[i])
	}
}

// MergeChains returns a single variable array from multiple chains (sampler/chain.go:96-148): a
// variable collapsed in ANY chain is reported by a clone of the first such chain's variable and takes
// no part in the summation; every other variable is the sum of the chains' Marginal vectors.  It
// only reads Target.Vars, which refresh() keeps current after every round.
func MergeChains(chains []*Chain) ([]*model.Variable, error) {
	if len(chains) < 1 {
		return nil, errors.Errorf("Can not merge 0 chains")
	}
	if len(chains) == 1 {
		return chains[0].Target.Vars, nil
	}
	n := len(chains[0].Target.Vars)
	merged := make([]*model.Variable, n)
	frozen := make([]bool, n)
	for i := 0; i < n; i++ {
		src := chains[0].Target.Vars[i]
		for _, ch := range chains {
			if i < len(ch.Target.Vars) && ch.Target.Vars[i].Collapsed {
				src, frozen[i] = ch.Target.Vars[i], true
				break
			}
		}
		merged[i] = src.Clone()
	}
	for _, ch := range chains[1:] {
		if len(ch.Target.Vars) != n {
			return nil, errors.Errorf("Cannot merge chain with %d vars into %d vars", len(ch.Target.Vars), n)
		}
		for i, v := range ch.Target.Vars {
			if frozen[i] {
				continue
			}
			for k, p := range v.Marginal {
				merged[i].Marginal[k] += p
			}
		}
	}
	return merged, nil
}

// ChainDist returns the (within-chain, between-chain) distance of one variable under distFunc
// (sampler/chain.go:253-290), averaged over this chain's replicas: each replica's half-window
// histograms (every bin seeded with 1e-8) give within = d(first half, second half) and
// between = d(merged, both halves).  ChainConvergence itself runs on the device; this host form
// exists for callers that hold a custom Measure.
func (c *Chain) ChainDist(distFunc Measure, varIdx int, mergedVar *model.Variable) (float64, float64, error) {
	if c.hist == nil {
		return math.NaN(), math.NaN(), errors.Errorf("Total seen %d < Convergence Window %d", 0, c.ConvergenceWindow)
	}
	vsrc := c.Target.Vars[varIdx]
	if vsrc.Card != mergedVar.Card {
		return math.NaN(), math.NaN(), errors.Errorf("Variable mismatch")
	}
	o, tc := 0, 0
	for i, v := range c.Target.Vars {
		if i < varIdx {
			o += v.Card
		}
		tc += v.Card
	}
	within, between := 0.0, 0.0
	for r := 0; r < c.replicas; r++ {
		v1, v2 := vsrc.Clone(), vsrc.Clone()
		for k := range vsrc.Marginal {
			v1.Marginal[k] = 1e-8 + float64(c.hist[(o+k)*c.replicas+r])
			v2.Marginal[k] = 1e-8 + float64(c.hist[(tc+o+k)*c.replicas+r])
		}
		within += distFunc(v1, v2)
		for k, p := range v2.Marginal {
			v1.Marginal[k] += p
		}
		between += distFunc(mergedVar, v1)
	}
	return within / float64(c.replicas), between / float64(c.replicas), nil
}

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
	runtime.LockOSThread() // keep gb_last_error on this thread
	defer runtime.UnlockOSThread()
	rc := C.gb_chains_convergence(thePool.h, measureID(distFunc), &merged[0], &out[0])
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
	rc = C.gb_chains_merged_marginals(thePool.h, &merged[0], nil)
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
