{
    no_decay: { opt+: { decay: 0 } },
}
