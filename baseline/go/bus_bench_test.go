// bus_bench_test.go — benchmark of the UNMODIFIED reference bus, for anyone with a Go toolchain.
//
// NOT RUN in this repository: the build image has no Go (SURVEY.md F1); the CPU numbers printed by bench.py come from
// oracle/gobus_baseline.c, a C restatement of the same cost model.  To run it, copy this file into the reference's
// events/ directory and:   go test -run xxx -bench BenchmarkPublish -benchtime 2s ./events/
package events

import (
	"fmt"
	"testing"
)

// one publisher, N subscribers with the production mailbox capacity (jobs/jobs.go:23), a consumer goroutine per
// subscriber draining its Rx exactly like Job.Run's select loop (jobs/jobs.go:173)
func benchmarkPublish(b *testing.B, n int) {
	bus := NewEventBus()
	subs := make([]*Subscriber, n)
	done := make(chan struct{})
	for i := range subs {
		s := &Subscriber{Rx: make(chan Event, 1000)}
		s.Subscribe(bus)
		subs[i] = s
		go func(s *Subscriber) {
			for {
				select {
				case <-s.Rx:
				case <-done:
					return
				}
			}
		}(s)
	}
	ev := Event{Code: StatusChanged, Source: "watch.backend"}
	b.ResetTimer()
	for i := 0; i < b.N; i++ {
		bus.Publish(ev)
	}
	b.StopTimer()
	close(done)
	b.ReportMetric(float64(b.N)*float64(n)/b.Elapsed().Seconds(), "deliveries/s")
}

func BenchmarkPublish(b *testing.B) {
	for _, n := range []int{8, 64, 1024, 8192, 65536} {
		b.Run(fmt.Sprintf("subs=%d", n), func(b *testing.B) { benchmarkPublish(b, n) })
	}
}
