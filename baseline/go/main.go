// baseline/go/main.go — run the UNMODIFIED reference (edp1096/toy-spice) on SPICE decks and dump its results.
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE (no Go toolchain, github.com/edp1096/sparse not vendored).  It exists so that
// anyone with Go can (a) turn this repo's "parity unpinned" into pinned parity and (b) measure the CPU baseline the
// north star asks for (the Go solver, one goroutine per instance, on the host cores).  Only the reference's public
// API is used, in the call order of cmd/spice/main.go:364-449.
//
//   cd <toy-spice checkout> && mkdir -p cmd/tsbdump && cp <this file> cmd/tsbdump/main.go
//   go run ./cmd/tsbdump -out <repo>/tests/golden/go  circuits/*.cir          # golden vectors (one JSON per deck)
//   go run ./cmd/tsbdump -bench -copies 4096 -workers 16 circuits/rlc.cir       # circuit-timesteps/s of the Go solver
//
// Golden vectors: tests/test_go_reference_vectors.py picks up tests/golden/go/<deck>.json automatically and holds the
// oracle (and, on a GPU box, the CUDA path) to them: row counts and TIME exactly, values at 1e-9 / 1e-12.
// JSON: {"deck": "...", "analysis": "tran|op|dc", "error": "", "results": {"TIME": [...], "V(1)": [...], ...}}
// (NaN / +-Inf are written as strings "NaN", "+Inf", "-Inf": encoding/json cannot represent them as numbers.)
package main

import (
	"encoding/json"
	"flag"
	"fmt"
	"log"
	"math"
	"os"
	"path/filepath"
	"strings"
	"sync"
	"time"

	"github.com/edp1096/toy-spice/pkg/analysis"
	"github.com/edp1096/toy-spice/pkg/circuit"
	"github.com/edp1096/toy-spice/pkg/netlist"
)

type dump struct {
	Deck     string                   `json:"deck"`
	Analysis string                   `json:"analysis"`
	Error    string                   `json:"error"`
	Results  map[string][]interface{} `json:"results"`
}

// build + run exactly as cmd/spice/main.go does; returns the analysis kind, GetResults() and Execute()'s error.
func runDeck(text string) (string, map[string][]float64, error) {
	nl, err := netlist.Parse(text)
	if err != nil {
		return "", nil, fmt.Errorf("parse: %v", err)
	}
	ckt := circuit.NewWithComplex(nl.Title, nl.Analysis == netlist.AnalysisAC)
	if err = ckt.AssignNodeBranchMaps(nl.Elements); err != nil {
		return "", nil, err
	}
	ckt.CreateMatrix()
	ckt.Models = nl.Models
	if err = ckt.SetupDevices(nl.Elements); err != nil {
		return "", nil, err
	}
	var an analysis.Analysis
	kind := ""
	switch nl.Analysis {
	case netlist.AnalysisOP:
		an, kind = analysis.NewOP(), "op"
	case netlist.AnalysisTRAN:
		p := nl.TranParam
		an, kind = analysis.NewTransient(p.TStart, p.TStop, p.TStep, p.TMax, p.UIC), "tran"
	case netlist.AnalysisDC:
		p := nl.DCParam
		an, kind = analysis.NewDCSweep([]string{p.Source1}, []float64{p.Start1}, []float64{p.Stop1}, []float64{p.Increment1}), "dc"
	default:
		return "", nil, fmt.Errorf("analysis not covered by the batched engine (AC)")
	}
	if err = an.Setup(ckt); err != nil {
		return kind, nil, err
	}
	err = an.Execute()
	return kind, an.GetResults(), err
}

func jsonable(v float64) interface{} {
	switch {
	case math.IsNaN(v):
		return "NaN"
	case math.IsInf(v, 1):
		return "+Inf"
	case math.IsInf(v, -1):
		return "-Inf"
	}
	return v
}

func main() {
	out := flag.String("out", ".", "directory for <deck>.json")
	bench := flag.Bool("bench", false, "time `copies` independent runs of each deck on `workers` goroutines")
	copies := flag.Int("copies", 1024, "instances per deck in -bench mode")
	workers := flag.Int("workers", 0, "goroutines in -bench mode (0 = GOMAXPROCS)")
	flag.Parse()
	for _, path := range flag.Args() {
		data, err := os.ReadFile(path)
		if err != nil {
			log.Fatal(err)
		}
		name := strings.TrimSuffix(filepath.Base(path), filepath.Ext(path))
		if *bench {
			// One goroutine per instance, as the north star words it.  Instances of one deck are identical here (the
			// reference has no parameter-override API); throughput of the solver does not depend on the values.
			// NOTE: decks with `core=` inductors share the package-level map netlist.magneticCores (parser.go:750):
			// run those with -workers 1.  bjt*.cir print from inside the device model (bjt.go) — redirect stdout.
			w := *workers
			if w <= 0 {
				w = 0
			}
			sem := make(chan struct{}, maxInt(1, w))
			if w == 0 {
				sem = make(chan struct{}, 1<<20)
			}
			var wg sync.WaitGroup
			var mu sync.Mutex
			steps := 0
			t0 := time.Now()
			for i := 0; i < *copies; i++ {
				wg.Add(1)
				sem <- struct{}{}
				go func() {
					defer wg.Done()
					defer func() { <-sem }()
					_, res, _ := runDeck(string(data))
					mu.Lock()
					steps += len(res["TIME"])
					mu.Unlock()
				}()
			}
			wg.Wait()
			dt := time.Since(t0).Seconds()
			fmt.Printf("{\"deck\": %q, \"copies\": %d, \"stored_rows\": %d, \"seconds\": %.6f, \"rows_per_sec\": %.6e}\n",
				name, *copies, steps, dt, float64(steps)/dt)
			continue
		}
		kind, res, runErr := runDeck(string(data))
		d := dump{Deck: name, Analysis: kind, Results: map[string][]interface{}{}}
		if runErr != nil {
			d.Error = runErr.Error()
		}
		for k, vals := range res {
			col := make([]interface{}, len(vals))
			for i, v := range vals {
				col[i] = jsonable(v)
			}
			d.Results[k] = col
		}
		f, err := os.Create(filepath.Join(*out, name+".json"))
		if err != nil {
			log.Fatal(err)
		}
		enc := json.NewEncoder(f)
		if err = enc.Encode(d); err != nil {
			log.Fatal(err)
		}
		f.Close()
		fmt.Printf("%s: %s, %d result series, error=%q\n", name, kind, len(res), d.Error)
	}
}

func maxInt(a, b int) int {
	if a > b {
		return a
	}
	return b
}
