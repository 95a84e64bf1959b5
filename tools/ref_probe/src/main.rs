//! Runs the programs of oracle/workloads.hpp (SLEEP, PINGPONG = BASELINE config 1, CHAOS) on the
//! REAL madsim executor and prints, per seed, two observables that pin the whole execution:
//!   elapsed_ns  virtual time at the main future's last poll (every latency draw and every
//!               50-100 ns poll advance before it is folded into this number)
//!   next_u64    the next output of the global RNG at that point (= how many draws were made)
//! `python expected.py --check probe.jsonl` compares them with what the oracle (and therefore the
//! device engine, which is bit-exact against the oracle) produces for the same seeds.
use madsim::net::Endpoint;
use madsim::rand::{random, thread_rng, Rng};
use madsim::runtime::{Handle, Runtime};
use madsim::time::{sleep, Instant};
use std::net::SocketAddr;
use std::sync::atomic::{AtomicU32, Ordering};
use std::sync::Arc;
use std::time::Duration;

const SEEDS: [u64; 6] = [0, 1, 2, 42, 2333, 4095];

fn probe(t0: Instant) -> (u64, u64) {
    (t0.elapsed().as_nanos() as u64, random::<u64>())
}

/// oracle/workloads.hpp wl_sleep_main
async fn wl_sleep() -> (u64, u64) {
    let t0 = Instant::now();
    sleep(Duration::from_secs(1)).await;
    probe(t0)
}

/// oracle/workloads.hpp wl_pingpong_main (endpoint.rs:550-584 repeated `rounds` times)
async fn wl_pingpong(rounds: u32) -> (u64, u64) {
    let t0 = Instant::now();
    let handle = Handle::current();
    let addr1: SocketAddr = "10.0.0.1:1".parse().unwrap();
    let node1 = handle.create_node().ip(addr1.ip()).build();
    let node2 = handle.create_node().ip("10.0.0.2".parse().unwrap()).build();
    node1.spawn(async move {
        let ep = Endpoint::bind(addr1).await.unwrap();
        assert_eq!(ep.local_addr().unwrap(), addr1);
        loop {
            let mut buf = vec![0u8; 16];
            let (len, from) = ep.recv_from(1, &mut buf).await.unwrap();
            assert_eq!(&buf[..len], b"ping");
            ep.send_to(from, 1, b"pong").await.unwrap();
        }
    });
    let f = node2.spawn(async move {
        sleep(Duration::from_secs(1)).await;
        let ep = Endpoint::connect(addr1).await.unwrap();
        assert_eq!(ep.peer_addr().unwrap(), addr1);
        let mut done = 0u32;
        for _ in 0..rounds {
            ep.send(1, b"ping").await.unwrap();
            let mut buf = vec![0u8; 16];
            let len = ep.recv(1, &mut buf).await.unwrap();
            assert_eq!(&buf[..len], b"pong");
            done += 1;
            sleep(Duration::from_millis(10)).await;
        }
        done
    });
    assert_eq!(f.await.unwrap(), rounds);
    probe(t0)
}

/// oracle/workloads.hpp wl_chaos_main (task/mod.rs:858-1016 kill / restart / pause_resume in one schedule)
async fn wl_chaos(rounds: u64, period: Duration) -> (u64, u64) {
    let t0 = Instant::now();
    let handle = Handle::current();
    let flag = Arc::new(AtomicU32::new(0));
    let starts = Arc::new(AtomicU32::new(0));
    let (flag1, starts1) = (flag.clone(), starts.clone());
    let node = handle
        .create_node()
        .init(move || {
            let (flag, starts) = (flag1.clone(), starts1.clone());
            async move {
                starts.fetch_add(1, Ordering::Relaxed);
                loop {
                    sleep(period).await;
                    flag.fetch_add(2, Ordering::Relaxed);
                }
            }
        })
        .build();
    let id = node.id();
    for _ in 0..rounds {
        let m: u64 = thread_rng().gen_range(1..4u64);
        sleep(period * (m as u32) + period / 2).await;
        let k: u64 = thread_rng().gen_range(0..3u64);
        if k == 0 {
            handle.kill(id);
            assert!(handle.is_exit(id));
            sleep(period).await;
            handle.restart(id);
            assert!(!handle.is_exit(id));
        } else if k == 1 {
            handle.pause(id);
            let before = flag.load(Ordering::Relaxed);
            sleep(period * 3).await;
            assert_eq!(flag.load(Ordering::Relaxed), before);
            handle.resume(id);
        } else {
            handle.restart(id);
        }
    }
    sleep(period).await;
    probe(t0)
}

fn main() {
    for seed in SEEDS {
        let run = |name: &str, out: (u64, u64)| {
            println!(
                "{{\"workload\":\"{name}\",\"seed\":{seed},\"elapsed_ns\":{},\"next_u64\":{}}}",
                out.0, out.1
            );
        };
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("sleep", rt.block_on(wl_sleep()));
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("pingpong", rt.block_on(wl_pingpong(1000)));
        let rt = Runtime::with_seed_and_config(seed, madsim::Config::default());
        run("chaos", rt.block_on(wl_chaos(8, Duration::from_millis(2000))));
    }
}
